mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mott" -s 2>&1 | tail -25
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/rnd2i_pytest_gpu.log; tail -5 gpurun_out/rnd2i_pytest_gpu.log
