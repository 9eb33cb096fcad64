mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "accumulated or default_tolerance or lookup" 2>&1 | tail -5
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/v3_guide2048.so build/ab/v3_guide4096.so > gpurun_out/rnd2d_ab_C2.jsonl 2>&1; cut -c1-160 gpurun_out/rnd2d_ab_C2.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/rnd2d_bench_2gpu.json 2> gpurun_out/rnd2d_bench_2gpu.err
echo "rc=$?"; tail -c 1500 gpurun_out/rnd2d_bench_2gpu.json; tail -5 gpurun_out/rnd2d_bench_2gpu.err
