mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -x -k "regrouping or tuning or pipelined or accumulated or error_behaviour or statistics_only" 2>&1 | tail -4
python scripts/ab_c3.py build/ab/new2.so build/ab/cur3.so build/ab/new2.so build/ab/cur3.so > gpurun_out/rnd2s_ab_c3.jsonl 2>&1; cut -c1-300 gpurun_out/rnd2s_ab_c3.jsonl
python scripts/ab_c3.py --cols type,energy --n 1000000 --workload Si build/ab/new2.so build/ab/cur3.so > gpurun_out/rnd2s_ab_c3_odd.jsonl 2>&1; cut -c1-300 gpurun_out/rnd2s_ab_c3_odd.jsonl
python scripts/ab_c3.py --cols type,position,pathl,E,E_loss,pol_angle,az_angle,energy,direction --n 400000 build/ab/new2.so build/ab/cur3.so > gpurun_out/rnd2s_ab_c3_all.jsonl 2>&1; cut -c1-300 gpurun_out/rnd2s_ab_c3_all.jsonl
python scripts/ab_kernel.py --workload C2 --pipes 4,6,8 MCD_TAIL_PRIORITY=0 MCD_TAIL_PRIORITY=1 MCD_TAIL_PRIORITY=0 MCD_TAIL_PRIORITY=1 > gpurun_out/rnd2s_ab_tailprio_C2.jsonl 2>&1; cut -c1-260 gpurun_out/rnd2s_ab_tailprio_C2.jsonl
for v in 0 1; do MCD_TAIL_PRIORITY=$v python bench.py --steps 20 --warmup 3 --legs none --no-cpu > gpurun_out/rnd2s_bench_tp$v.json 2>/dev/null; done
python - <<'PY'
import json
for f in ('tp0','tp1'):
    for l in open('gpurun_out/rnd2s_bench_%s.json' % f):
        if l.startswith('{') and '"metric"' in l:
            d = json.loads(l)
            print(f, 'value %.4g ms/step %.3f e2e %.4g ratio %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value']/d['value']), d.get('e2e_detail'))
PY
