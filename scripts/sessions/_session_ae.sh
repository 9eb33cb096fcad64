bash scripts/gpu_session.sh rnd2ae tests --steps 20 --warmup 3
bash scripts/ncu_capture.sh rnd2ae
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_paths.py > gpurun_out/rnd2ae_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/rnd2ae_sanitizer_memcheck.log
