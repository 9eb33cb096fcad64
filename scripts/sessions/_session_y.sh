bash scripts/gpu_session.sh rnd2y tests --steps 20 --warmup 3
bash scripts/ncu_capture.sh rnd2y
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_paths.py > gpurun_out/rnd2y_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/rnd2y_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_paths.py > gpurun_out/rnd2y_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/rnd2y_sanitizer_racecheck.log
