mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gsweep_kernel -s 1 -c 1 -f -o gpurun_out/r02c_gsweep_c3 python tools/profile_sweep.py c3 151552 > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gsweep_kernel -s 1 -c 1 -f -o gpurun_out/r02c_gsweep_c5 python tools/profile_gsweep.py 284160 10000 > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out/r02b*
