mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_paths.py > gpurun_out/rnd2q_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/rnd2q_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_paths.py > gpurun_out/rnd2q_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/rnd2q_sanitizer_racecheck.log
bash scripts/gpu_session.sh rnd2q tests --steps 20 --warmup 3
bash scripts/ncu_capture.sh rnd2q
