mkdir -p gpurun_out
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/v1_lookup.so build/ab/v1_wsstream.so build/ab/v1_guide2048.so build/ab/v1_slots56.so build/ab/v1_lookup.so > gpurun_out/rnd2c_ab_C2.jsonl 2>&1
cut -c1-200 gpurun_out/rnd2c_ab_C2.jsonl
bash scripts/ncu_capture.sh rnd2c
