"""Summarise an ncu report per CUDA source line: share of executed warp instructions and of stall samples.
usage: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = ''; hdr = None; lines = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0]:       # a CUDA line row (SASS rows have empty line no)
        lines.append((cur_file, r))
ix = {h: i for i, h in enumerate(hdr)}
def num(x):
    try: return float(x)
    except Exception: return 0.0
I = hdr.index('Instructions Executed'); S = hdr.index('# Samples'); L = hdr.index('stall_long_sb')
T = hdr.index('Thread Instructions Executed')
ti = sum(num(r[I]) for _, r in lines); ts = sum(num(r[S]) for _, r in lines)
print('total warp instructions %.4g, samples %d' % (ti, ts))
lines.sort(key=lambda fr: -num(fr[1][I if (len(sys.argv) > 3 and sys.argv[3] == "inst") else S]))
print('%7s %7s %7s %5s  %s' % ('inst%', 'samp%', 'longsb%', 'thr', 'line'))
for f, r in lines[:top]:
    print('%6.2f%% %6.2f%% %6.2f%% %5.1f  %s:%s  %s' % (100 * num(r[I]) / ti, 100 * num(r[S]) / ts, 100 * num(r[L]) / max(num(r[S]), 1),
          num(r[T]) / max(num(r[I]), 1), f, r[0], r[1].strip()[:95]))
