#!/bin/bash
# ncu evidence for the trajectory kernel (run through gpurun on ONE GPU):
#   1. launch list of the headline bench command      -> gpurun_out/<tag>_launches.csv
#   2. ncu --set full of the passes of one C2 step     -> gpurun_out/<tag>_pool.ncu-rep (+ raw / lines / opcodes summaries)
#   3. profiles/traj_kernel_summary.json (per-event instruction count and DRAM traffic for bench.py)
tag=${1:-p}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --legs none --no-cpu > gpurun_out/${tag}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traj_kernel_pool -s 3 -c 3 -f -o gpurun_out/${tag}_pool \
  python scripts/run_configs.py C2 > gpurun_out/${tag}_full.log 2>&1
python scripts/ncu_raw.py gpurun_out/${tag}_pool.ncu-rep gpurun_out/${tag}_traj_kernel_ncu.json "ncu --set full, the three launches of one C2 step (${tag})" > gpurun_out/${tag}_raw.txt 2>&1
python scripts/ncu_lines.py gpurun_out/${tag}_pool.ncu-rep 60 > gpurun_out/${tag}_traj_lines.txt 2>&1
python scripts/ncu_opcodes.py gpurun_out/${tag}_pool.ncu-rep > gpurun_out/${tag}_traj_opcodes.txt 2>&1
ev=$(python - <<PY
import json
for l in open('gpurun_out/${tag}_full.log'):
    if l.startswith('{'):
        d = json.loads(l); print(d['events_per_electron'] * 1e6); break
PY
)
python scripts/ncu_summary.py ${tag} gpurun_out/${tag}_traj_kernel_ncu.json $ev gpurun_out/${tag}_traj_kernel_summary.json
tail -3 gpurun_out/${tag}_raw.txt
