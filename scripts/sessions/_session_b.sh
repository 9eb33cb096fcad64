mkdir -p gpurun_out
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/base.so build/ab/v1_lookup.so build/ab/v2_sub1.so build/ab/v2_sub2.so build/ab/v2_sub3.so > gpurun_out/rnd2b_ab_C2.jsonl 2>&1
python scripts/ab_kernel.py --workload Au --pipes 6 --n 500000 build/ab/base.so build/ab/v1_lookup.so build/ab/v2_sub2.so build/ab/v2_sub3.so > gpurun_out/rnd2b_ab_Au.jsonl 2>&1
python scripts/ab_kernel.py --workload Cu --pipes 6 build/ab/base.so build/ab/v2_sub2.so > gpurun_out/rnd2b_ab_Cu.jsonl 2>&1
cat gpurun_out/rnd2b_ab_*.jsonl | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -120 > gpurun_out/rnd2b_pytest_gpu.log; tail -5 gpurun_out/rnd2b_pytest_gpu.log
python scripts/diag_c5_cell.py Al 5000 cont JL > gpurun_out/rnd2b_diag_c5.log 2>&1; tail -8 gpurun_out/rnd2b_diag_c5.log
