mkdir -p gpurun_out
# ncu --set full of the continuous kernel (C1 shape at 1e6 electrons, classical and explicit Bethe) and of the row-recording kernel + regrouping (C3 batch)
cat > /tmp/cont_run.py <<'PY'
import sys; sys.path.insert(0, '.')
from mcdiscrete_b200.engine import Engine
m, model = sys.argv[1], sys.argv[2]
ip = dict(material=m, E0=20000., Emin=500., mode='cont', Bethe_model=model, s_tilt=70., maxScatt=100000, thickness=3e6, Wc=10.,
          electron_output=['outcome', 'final_E', 'final_dir', 'last_pos'], scatter_output=[], seed=12345)
eng = Engine(ip, device=0)
for rep in range(3):
    r = eng.run(count=1000000, first_id=rep * 1000000, drain=False)
    print(m, model, r.kernel_ms, r.counters['events'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traj_kernel -s 2 -c 1 -f -o gpurun_out/rnd2af_cont_cl python /tmp/cont_run.py Al classical > gpurun_out/rnd2af_cont_cl.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traj_kernel -s 2 -c 1 -f -o gpurun_out/rnd2af_cont_ex python /tmp/cont_run.py Cu explicit > gpurun_out/rnd2af_cont_ex.log 2>&1
python scripts/ncu_raw.py gpurun_out/rnd2af_cont_cl.ncu-rep gpurun_out/rnd2af_cont_cl_ncu.json "ncu --set full, Al continuous classical, 1e6 electrons" > gpurun_out/rnd2af_cont_cl_raw.txt 2>&1
python scripts/ncu_raw.py gpurun_out/rnd2af_cont_ex.ncu-rep gpurun_out/rnd2af_cont_ex_ncu.json "ncu --set full, Cu continuous explicit, 1e6 electrons" > gpurun_out/rnd2af_cont_ex_raw.txt 2>&1
python scripts/ncu_lines.py gpurun_out/rnd2af_cont_cl.ncu-rep 50 > gpurun_out/rnd2af_cont_cl_lines.txt 2>&1
grep -h "^Al\|^Cu" gpurun_out/rnd2af_cont_cl.log gpurun_out/rnd2af_cont_ex.log | tail -4
grep -E "time_duration|inst_executed.sum|issue_active|per_inst_executed|pipe_xu|pipe_fma|pipe_alu|registers" gpurun_out/rnd2af_cont_cl_raw.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"traj_kernel|scatter_chunks" -s 6 -c 3 -f -o gpurun_out/rnd2af_c3 python scripts/profile_c3.py > gpurun_out/rnd2af_c3.log 2>&1
python scripts/ncu_raw.py gpurun_out/rnd2af_c3.ncu-rep gpurun_out/rnd2af_c3_ncu.json "ncu --set full, C3 batch: row kernel + regrouping" > gpurun_out/rnd2af_c3_raw.txt 2>&1
tail -4 gpurun_out/rnd2af_c3_raw.txt; tail -5 gpurun_out/rnd2af_c3.log
