bash scripts/gpu_session.sh rnd2f tests --steps 20 --warmup 3
bash scripts/ncu_capture.sh rnd2f
