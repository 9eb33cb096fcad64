mkdir -p gpurun_out
timeout 1500 python scripts/stat_report.py --default-tol > gpurun_out/rnd2v_statistics_default_tol.jsonl 2> gpurun_out/rnd2v_statistics.err; echo "rc=$?"
cut -c1-420 gpurun_out/rnd2v_statistics_default_tol.jsonl; tail -3 gpurun_out/rnd2v_statistics.err
