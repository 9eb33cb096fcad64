mkdir -p gpurun_out
python scripts/ab_kernel.py --workload C2 --pipes 6 build/ab/cur3.so build/ab/redux.so build/ab/redux_bfast.so build/ab/cur3.so build/ab/redux.so build/ab/redux_bfast.so > gpurun_out/rnd2u_ab_C2.jsonl 2>&1; cut -c1-300 gpurun_out/rnd2u_ab_C2.jsonl
python scripts/ab_kernel.py --workload Au --pipes 6 --n 500000 build/ab/cur3.so build/ab/redux.so build/ab/redux_bfast.so > gpurun_out/rnd2u_ab_Au.jsonl 2>&1; cut -c1-300 gpurun_out/rnd2u_ab_Au.jsonl
