mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -x -k "mott or regrouping or error_behaviour" 2>&1 | tail -5
python scripts/ab_c3.py build/ab/head.so build/ab/new.so build/ab/new_t256.so build/ab/new_t128.so > gpurun_out/rnd2j_ab_c3.jsonl 2>&1; cut -c1-330 gpurun_out/rnd2j_ab_c3.jsonl
python scripts/ab_c3.py --cols type,position,pathl,E,E_loss,pol_angle,az_angle,energy,direction --n 500000 build/ab/head.so build/ab/new.so > gpurun_out/rnd2j_ab_c3_allcols.jsonl 2>&1; cut -c1-330 gpurun_out/rnd2j_ab_c3_allcols.jsonl
python scripts/ab_kernel.py --workload cont --pipes 2 build/ab/new.so build/ab/new_t256.so build/ab/new_t128.so > gpurun_out/rnd2j_ab_cont.jsonl 2>&1; cut -c1-200 gpurun_out/rnd2j_ab_cont.jsonl
python scripts/ab_kernel.py --workload contEx --pipes 2 build/ab/new.so build/ab/new_t256.so > gpurun_out/rnd2j_ab_contEx.jsonl 2>&1; cut -c1-200 gpurun_out/rnd2j_ab_contEx.jsonl
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/head.so build/ab/new.so > gpurun_out/rnd2j_ab_C2.jsonl 2>&1; cut -c1-200 gpurun_out/rnd2j_ab_C2.jsonl
