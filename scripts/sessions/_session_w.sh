mkdir -p gpurun_out
# how much of the step is exposed table latency?  the same kernel on tables small enough to live in L2 (tol 1e-2 / 1e-3)
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/final.so > gpurun_out/rnd2w_C2_default_tables.jsonl 2>&1; cut -c1-200 gpurun_out/rnd2w_C2_default_tables.jsonl
python scripts/ab_kernel.py --workload C2 --pipes 6 --tol 1e-3,1e-5 build/ab/final.so > gpurun_out/rnd2w_C2_mid_tables.jsonl 2>&1; cut -c1-200 gpurun_out/rnd2w_C2_mid_tables.jsonl
python scripts/ab_kernel.py --workload C2 --pipes 6 --tol 1e-2,1e-3 build/ab/final.so > gpurun_out/rnd2w_C2_small_tables.jsonl 2>&1; cut -c1-200 gpurun_out/rnd2w_C2_small_tables.jsonl
