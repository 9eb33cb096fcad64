mkdir -p gpurun_out
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/head.so build/ab/tk1.so build/ab/tk8.so build/ab/tk16.so build/ab/tk32.so build/ab/tk32r8.so build/ab/head.so > gpurun_out/rnd2g_ab_C2.jsonl 2>&1
cut -c1-260 gpurun_out/rnd2g_ab_C2.jsonl
python scripts/ab_kernel.py --workload Au --pipes 6 --n 500000 build/ab/head.so build/ab/tk8.so build/ab/tk32.so > gpurun_out/rnd2g_ab_Au.jsonl 2>&1
cut -c1-260 gpurun_out/rnd2g_ab_Au.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cumulative_cross or statistics_only" 2>&1 | tail -5
