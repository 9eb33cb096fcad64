"""Executed warp-instruction histogram by SASS opcode from an ncu report.  usage: ncu_opcodes.py rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
I = hdr.index('Instructions Executed'); S = hdr.index('Source'); T = hdr.index('Thread Instructions Executed')
cnt = collections.Counter(); thr = collections.Counter()
for r in rows[2:]:
    if len(r) != len(hdr): continue
    src = r[S].strip()
    if src.startswith('@'): src = src.split(None, 1)[1]
    op = src.split()[0].split('.')[0]
    if src.split()[0].startswith('MUFU'): op = src.split()[0]
    try:
        cnt[op] += float(r[I]); thr[op] += float(r[T])
    except ValueError: pass
tot = sum(cnt.values())
print('total warp instructions %.4g' % tot)
for op, c in cnt.most_common(top):
    print('%-14s %6.2f%%  avg lanes %5.1f' % (op, 100 * c / tot, thr[op] / max(c, 1)))
