mkdir -p gpurun_out
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/head.so build/ab/t32b32.so build/ab/t64b16.so build/ab/t128b8.so build/ab/head.so > gpurun_out/rnd2h_ab_C2.jsonl 2>&1
cut -c1-260 gpurun_out/rnd2h_ab_C2.jsonl
python scripts/ab_kernel.py --workload Au --pipes 6 --n 500000 build/ab/head.so build/ab/t32b32.so build/ab/t64b16.so > gpurun_out/rnd2h_ab_Au.jsonl 2>&1
cut -c1-260 gpurun_out/rnd2h_ab_Au.jsonl
