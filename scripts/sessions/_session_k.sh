mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -x -k "mott or regrouping or error_behaviour or tuning or accumulated or pipelined" 2>&1 | tail -5
python scripts/ab_c3.py build/ab/new.so build/ab/new2.so > gpurun_out/rnd2k_ab_c3.jsonl 2>&1; cut -c1-330 gpurun_out/rnd2k_ab_c3.jsonl
python scripts/ab_c3.py --cols type,energy --n 1000000 --workload Si build/ab/new.so build/ab/new2.so > gpurun_out/rnd2k_ab_c3_odd.jsonl 2>&1; cut -c1-330 gpurun_out/rnd2k_ab_c3_odd.jsonl
