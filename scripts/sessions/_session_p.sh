mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -x -k "pipelined or accumulated or statistics_only or tuning or run_properties" 2>&1 | tail -4
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) "$@"; }
run 4 bench.py --gpus 4 --steps 20 --warmup 3 --legs none --no-cpu > gpurun_out/rnd2p_bench_4gpu.json 2> gpurun_out/rnd2p_bench_4gpu.err; echo "rc=$?"
python bench.py --steps 20 --warmup 3 --legs none --no-cpu > gpurun_out/rnd2p_bench_1gpu.json 2> gpurun_out/rnd2p_bench_1gpu.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('4gpu','1gpu'):
    for l in open('gpurun_out/rnd2p_bench_%s.json' % f):
        if l.startswith('{') and '"metric"' in l:
            d = json.loads(l)
            print(f, 'value %.4g ms/step %.3f e2e %.4g ratio %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value']/d['value']), d.get('e2e_detail'))
PY
