# 8-GPU session: north-star bench at N=8 (strong-scaling legs + checksums), C5 sweep at 8 / 4 / 2 GPUs
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/rnd2m_gpus.txt 2>&1
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
run 8 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/rnd2m_bench_8gpu.json 2> gpurun_out/rnd2m_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 900 gpurun_out/rnd2m_bench_8gpu.json
for n in 8 4 2; do
  run $n scripts/run_c5.py > gpurun_out/rnd2m_c5_${n}gpu.jsonl 2> gpurun_out/rnd2m_c5_${n}gpu.err; echo "c5 x$n rc=$?"; tail -1 gpurun_out/rnd2m_c5_${n}gpu.jsonl
done
run 4 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/rnd2m_bench_4gpu.json 2> gpurun_out/rnd2m_bench_4gpu.err; echo "bench4 rc=$?"; tail -c 400 gpurun_out/rnd2m_bench_4gpu.json
