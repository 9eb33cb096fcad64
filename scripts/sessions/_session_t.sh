mkdir -p gpurun_out
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) "$@"; }
run 2 scripts/check_multigpu.py > gpurun_out/rnd2t_check_multigpu.log 2>&1; echo "check_multigpu rc=$?"; tail -3 gpurun_out/rnd2t_check_multigpu.log
run 8 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/rnd2t_bench_8gpu.json 2> gpurun_out/rnd2t_bench_8gpu.err; echo "bench8 rc=$?"
run 2 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/rnd2t_bench_2gpu.json 2> gpurun_out/rnd2t_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ('8gpu','2gpu'):
    for l in open('gpurun_out/rnd2t_bench_%s.json' % f):
        if l.startswith('{') and '"metric"' in l:
            d = json.loads(l)
            L = d['legs']
            print(f, 'value %.4g ms/step %.3f e2e %.4g ratio %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value']/d['value']), d.get('e2e_detail'))
            print('   target %.3f s %s c4 %.3f s %s full %.3g rows/s %.1f GB/s %s c5 %s' % (L['target']['wall_s'], L['target']['stats_sha1'], L['c4']['wall_s'], L['c4']['stats_sha1'], L['full_output']['rows_per_s'], L['full_output']['d2h_GBps'], L['full_output']['stats_sha1'], L['c5_cells']['sha1_of_digests']))
PY
