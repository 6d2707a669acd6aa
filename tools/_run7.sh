mkdir -p gpurun_out
python tools/profile_sweep.py c3 151552 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gsweep_kernel -s 1 -c 1 -f -o gpurun_out/gsweep_c3 python tools/profile_sweep.py c3 151552 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
