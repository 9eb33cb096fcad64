"""profiles/traj_kernel_summary.json from an `ncu --set full` capture of the passes of one C2 step (scripts/ncu_raw.py
output) -- what bench.py reads for `roofline.traffic` and the instruction-issue roof.

usage: python scripts/ncu_summary.py <tag> <ncu_raw.json> <events per step> [out.json]"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, raw, events = sys.argv[1], json.load(open(sys.argv[2])), float(sys.argv[3])
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, 'profiles', 'traj_kernel_summary.json')
h = hashlib.sha1()
for f in ('mcd_traj.cu', 'mcd_device.cuh'):
    h.update(open(os.path.join(ROOT, 'mcdiscrete_b200', 'csrc', f), 'rb').read())
launches = raw.get('launches', [raw])
def num(l, k):
    try: return float(l[k]['value'])
    except Exception: return 0.0
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
dram = sum(num(l, k) * UNIT.get(l[k]['unit'], 1.0) for l in launches for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum') if k in l)
inst = sum(num(l, 'smsp__inst_executed.sum') for l in launches)
main = launches[0]
d = {'tag': tag, 'kernel': 'traj_kernel_pool<3> (main pass + %d tail hand-over passes = one C2 step of 1e6 electrons)' % (len(launches) - 1),
     'kernel_source_sha1': h.hexdigest()[:16], 'events_per_step': events,
     'warp_inst_per_step': inst, 'warp_inst_per_event': inst / events,
     'dram_bytes_per_step': dram, 'dram_bytes_per_event': dram / events,
     'per_pass_ms': [num(l, 'gpu__time_duration.sum') for l in launches],
     'main_pass': {'issue_active_pct': num(main, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                   'lanes_per_inst': num(main, 'smsp__thread_inst_executed_per_inst_executed.ratio'),
                   'xu_pct': num(main, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
                   'fma_pct': num(main, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
                   'alu_pct': num(main, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
                   'l1_data_pipe_pct': num(main, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
                   'l1_hit_pct': num(main, 'l1tex__t_sector_hit_rate.pct'), 'l2_hit_pct': num(main, 'lts__t_sector_hit_rate.pct'),
                   'dram_pct': num(main, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                   'regs': num(main, 'launch__registers_per_thread')},
     'source': 'ncu --set full --clock-control none, the launches of one step launched alone (cold-cache, serialised)'}
json.dump(d, open(out, 'w'), indent=1)
print(json.dumps(d))
