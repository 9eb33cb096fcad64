"""Print the headline raw metrics of an ncu report (and optionally dump them as JSON).  A report that holds several
kernel launches (the passes of one run) is dumped as {"launches": [...], "sum": {...}}.
usage: python scripts/ncu_raw.py report.ncu-rep [out.json] [note]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
more = rows[3:]
keys = ['Kernel Name', 'launch__grid_size', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'sm__cycles_elapsed.max']
out = {k: {'value': d[k][0], 'unit': d[k][1]} for k in keys if k in d}
for k, v in out.items():
    print('%-85s %s %s' % (k, v['value'], v['unit']))
if more:
    launches = [out]
    for r in more:
        dd = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        launches.append({k: {'value': dd[k][0], 'unit': dd[k][1]} for k in keys if k in dd})
    def num(x):
        try: return float(x)
        except Exception: return 0.0
    tot = {k: sum(num(l[k]['value']) for l in launches) for k in ('gpu__time_duration.sum', 'smsp__inst_executed.sum')}
    UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot['dram_bytes'] = sum(num(l[k]['value']) * UNIT.get(l[k]['unit'], 1.0) for l in launches
                            for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    for i, l in enumerate(launches[1:], 1):
        print('launch %d: %s ms, %s inst, issue %s %%' % (i, l['gpu__time_duration.sum']['value'], l['smsp__inst_executed.sum']['value'],
              l['smsp__issue_active.avg.pct_of_peak_sustained_active']['value']))
    print('sum', tot)
    out = {'launches': launches, 'sum': tot}
if len(sys.argv) > 2:
    if len(sys.argv) > 3:
        out['note'] = sys.argv[3]
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
