"""A/B timing of trajectory-kernel build variants (tuning tool, not part of the product).

    python scripts/ab_kernel.py [--workload C2|Cu|Au|cont] [--n N] [--reps R] [--pipes 2,3,4] variant ...

A variant is a library path (lib1.so) or environment settings of the tuning knobs (MCD_DUMP_T=16,MCD_PASSES=2).

Every library is loaded in its own process (MCD_LIB), runs the same electron ids, and prints the kernel time
(CUDA events inside mcd_run), the pipelined two-stream step time and a checksum of the statistics block; the
checksums of all variants must be equal (the result does not depend on lane scheduling)."""
import argparse, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WL = {
    'C2': dict(material='Si', E0=20000., Emin=10000., mode='DS', s_tilt=70.),
    'Cu': dict(material='Cu', E0=20000., Emin=10000., mode='DS', s_tilt=70.),
    'Au': dict(material='Au', E0=30000., Emin=15000., mode='DS', s_tilt=0., maxScatt=10000),
    'cont': dict(material='Al', E0=20000., Emin=500., mode='continuous', Bethe_model='classical', s_tilt=70.),
    'contJL': dict(material='Al', E0=20000., Emin=500., mode='continuous', Bethe_model='JL', s_tilt=70.),
    'contEx': dict(material='Cu', E0=20000., Emin=500., mode='continuous', Bethe_model='explicit', s_tilt=70.),
    'contAuEx': dict(material='Au', E0=30000., Emin=1500., mode='continuous', Bethe_model='explicit', s_tilt=70., maxScatt=100000),
}


def child(a):
    import torch
    from mcdiscrete_b200.engine import Engine
    from mcdiscrete_b200.probTables import genTables
    ip = dict(maxScatt=1000, thickness=3e6, Wc=10., tol_E=1e-4, tol_W=1e-7,
              electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], scatter_output=[], seed=12345)
    ip.update(WL[a.workload])
    if a.tol:
        ip['tol_E'], ip['tol_W'] = [float(x) for x in a.tol.split(',')]
    eng = Engine(ip, device=0)
    if ip['mode'] == 'DS':
        genTables(ip, engine=eng)
    eng.set_capacity(a.n, 0)
    st = torch.cuda.current_stream().cuda_stream
    ks = []
    for rep in range(a.reps + 2):
        eng.reset_stats(st)
        eng.launch(rep * a.n, a.n, st)
        eng.sync(st)
        if rep >= 2:
            ks.append(eng.kernel_ms())
    eng.reset_stats(st)
    eng.launch(0, a.n, st)
    eng.sync(st)
    c = eng.counters()
    h = hashlib.sha1(b''.join(v[0].tobytes() for v in eng.histograms().values())).hexdigest()[:12]
    # pipelined: two contexts / two streams like bench.py
    engs = [eng] + [eng.clone() for _ in range(max(a.pipes) - 1)]
    for e in engs[1:]:
        e.set_capacity(a.n, 0)
    ts = [torch.cuda.Stream() for _ in engs]
    piped = {}
    for np_ in a.pipes:
        def step(k):
            with torch.cuda.stream(ts[k % np_]):
                engs[k % np_].reset_stats(ts[k % np_].cuda_stream)
                engs[k % np_].launch((100 + k) * a.n, a.n, ts[k % np_].cuda_stream)
        for k in range(2 * np_):
            step(k)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = [torch.cuda.Event(enable_timing=True) for _ in range(np_)]
        e0.record(ts[0])
        for t in ts[1:np_]:
            t.wait_event(e0)
        nst = a.reps * 4
        for k in range(nst):
            step(k)
        for i in range(np_):
            e1[i].record(ts[i])
        torch.cuda.synchronize()
        piped[np_] = round(max(e0.elapsed_time(x) for x in e1) / nst, 4)
    ms = min(piped.values())
    print(json.dumps({'variant': a.tag, 'kernel_ms': sorted(ks)[len(ks) // 2],
                      'kernel_ms_min': min(ks), 'pipelined_ms': piped, 'el_per_s': a.n / ms * 1e3,
                      'events': c['events'], 'outcomes': c['outcome'], 'hist_sha': h, 'launch': eng.launch_info()}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tol', default='', help='tol_E,tol_W (default: BASELINE defaults 1e-4,1e-7)')
    ap.add_argument('--workload', default='C2')
    ap.add_argument('--n', type=int, default=1000000)
    ap.add_argument('--reps', type=int, default=8)
    ap.add_argument('--child', action='store_true')
    ap.add_argument('--tag', default='')
    ap.add_argument('--pipes', type=lambda v: [int(x) for x in v.split(',')], default=[2, 3, 4])
    ap.add_argument('variants', nargs='*', help='a library path, or KEY=VAL[,KEY=VAL...] environment settings')
    a = ap.parse_args()
    if a.child:
        return child(a)
    for var in a.variants or ['default=1']:
        env = dict(os.environ)
        if var.endswith('.so'):
            env['MCD_LIB'] = os.path.abspath(var)
        else:
            env.update(dict(kv.split('=', 1) for kv in var.split(',')))
        cmd = [sys.executable, os.path.abspath(__file__), '--child', '--workload', a.workload, '--n', str(a.n),
               '--reps', str(a.reps), '--tag', var, '--pipes', ','.join(map(str, a.pipes))] + (['--tol', a.tol] if a.tol else [])
        try:      # a variant that hangs must not eat the GPU budget
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=120)
            sys.stdout.write(r.stdout if r.returncode == 0 else 'FAILED %s\n%s\n' % (var, r.stderr[-2000:]))
        except subprocess.TimeoutExpired:
            sys.stdout.write('TIMEOUT %s\n' % var)
        sys.stdout.flush()


if __name__ == '__main__':
    main()
