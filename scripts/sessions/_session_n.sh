mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|^CPU(s)\|model name"; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))") > gpurun_out/rnd2n_topo.txt 2>&1
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
MCD_NO_BIND=1 run 4 bench.py --gpus 4 --steps 40 --warmup 3 --legs full_output --no-cpu > gpurun_out/rnd2n_bench_4gpu_nobind.json 2> gpurun_out/rnd2n_bench_4gpu_nobind.err; echo "nobind rc=$?"
run 4 bench.py --gpus 4 --steps 40 --warmup 3 --legs full_output --no-cpu > gpurun_out/rnd2n_bench_4gpu_bind.json 2> gpurun_out/rnd2n_bench_4gpu_bind.err; echo "bind rc=$?"
python - <<'PY'
import json
for f in ('nobind','bind'):
    for l in open('gpurun_out/rnd2n_bench_4gpu_%s.json' % f):
        if l.startswith('{') and '"metric"' in l:
            d = json.loads(l)
            print(f, 'value %.4g e2e %.4g d2h %.1f GB/s rows/s %.3g' % (d['value'], d['e2e']['value'], d['legs']['full_output']['d2h_GBps'], d['legs']['full_output']['rows_per_s']), d.get('host_binding'))
PY
python scripts/run_c5.py > gpurun_out/rnd2n_c5_1gpu.jsonl 2> gpurun_out/rnd2n_c5_1gpu.err; tail -1 gpurun_out/rnd2n_c5_1gpu.jsonl
cat gpurun_out/rnd2n_topo.txt | head -30
