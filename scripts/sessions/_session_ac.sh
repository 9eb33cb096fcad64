mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "cont or stopping or explicit or golden_formula or reference_distributions" 2>&1 | tail -4
python scripts/ab_kernel.py --workload cont --pipes 6 build/ab/onelog.so build/ab/nodiv.so build/ab/onelog.so build/ab/nodiv.so > gpurun_out/rnd2ac_ab_cont.jsonl 2>&1; cut -c1-330 gpurun_out/rnd2ac_ab_cont.jsonl
python scripts/ab_kernel.py --workload contJL --pipes 6 build/ab/onelog.so build/ab/nodiv.so > gpurun_out/rnd2ac_ab_contJL.jsonl 2>&1; cut -c1-330 gpurun_out/rnd2ac_ab_contJL.jsonl
python scripts/ab_kernel.py --workload contEx --pipes 6 build/ab/onelog.so build/ab/nodiv.so > gpurun_out/rnd2ac_ab_contEx.jsonl 2>&1; cut -c1-330 gpurun_out/rnd2ac_ab_contEx.jsonl
