# final numerics (division-free Bethe): digests at N = 1, 2, 4, 8 once more
mkdir -p gpurun_out
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((30000 + n)) "$@"; }
for n in 8 4 2; do
  run $n scripts/run_c5.py > gpurun_out/rnd2ad_c5_${n}gpu.jsonl 2> gpurun_out/rnd2ad_c5_${n}gpu.err; echo "c5 x$n rc=$?"; tail -1 gpurun_out/rnd2ad_c5_${n}gpu.jsonl
done
python scripts/run_c5.py > gpurun_out/rnd2ad_c5_1gpu.jsonl 2> gpurun_out/rnd2ad_c5_1gpu.err; echo "c5 x1 rc=$?"; tail -1 gpurun_out/rnd2ad_c5_1gpu.jsonl
run 8 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/rnd2ad_bench_8gpu.json 2> gpurun_out/rnd2ad_bench_8gpu.err; echo "bench8 rc=$?"
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/rnd2ad_bench_1gpu.json 2> gpurun_out/rnd2ad_bench_1gpu.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for f in ('8gpu','1gpu'):
    for l in open('gpurun_out/rnd2ad_bench_%s.json' % f):
        if l.startswith('{') and '"metric"' in l:
            d = json.loads(l); L = d['legs']
            print(f, 'value %.4g ms/step %.3f e2e %.4g ratio %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value']/d['value']))
            print('   target %.3f s %s c4 %.3f s %s full %.3g rows/s %.1f GB/s %s c5 %s' % (L['target']['wall_s'], L['target']['stats_sha1'], L['c4']['wall_s'], L['c4']['stats_sha1'], L['full_output']['rows_per_s'], L['full_output']['d2h_GBps'], L['full_output']['stats_sha1'], L['c5_cells']['sha1_of_digests']))
PY
