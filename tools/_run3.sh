set -x
python tools/host_profile.py config2 > gpurun_out/hostprof_cfg2_resident.log 2>&1
python tools/host_profile_e2e.py config2 > gpurun_out/hostprof_cfg2_e2e.log 2>&1
python tools/host_profile_e2e.py config1 > gpurun_out/hostprof_cfg1_e2e.log 2>&1
head -60 gpurun_out/hostprof_cfg2_e2e.log
python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_cfg3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/profile_trace.py 2097152 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_intersect3d -c 1 -f -o gpurun_out/prof_r1_hybrid python tools/profile_trace.py 2097152 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
