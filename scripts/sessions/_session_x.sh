mkdir -p gpurun_out
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/final.so build/ab/t192b6.so build/ab/b5.so build/ab/t288b4.so build/ab/final.so build/ab/t192b6.so > gpurun_out/rnd2x_ab_occupancy_C2.jsonl 2>&1; cut -c1-200 gpurun_out/rnd2x_ab_occupancy_C2.jsonl
python scripts/ab_kernel.py --workload Au --pipes 6 --n 500000 build/ab/final.so build/ab/t192b6.so build/ab/b5.so > gpurun_out/rnd2x_ab_occupancy_Au.jsonl 2>&1; cut -c1-200 gpurun_out/rnd2x_ab_occupancy_Au.jsonl
