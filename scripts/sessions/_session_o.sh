mkdir -p gpurun_out
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/cur.so build/ab/parktake.so build/ab/cur.so build/ab/parktake.so > gpurun_out/rnd2o_ab_C2.jsonl 2>&1; cut -c1-330 gpurun_out/rnd2o_ab_C2.jsonl
python scripts/ab_kernel.py --workload Au --pipes 6 --n 500000 build/ab/cur.so build/ab/parktake.so > gpurun_out/rnd2o_ab_Au.jsonl 2>&1; cut -c1-330 gpurun_out/rnd2o_ab_Au.jsonl
