mkdir -p gpurun_out
python tools/profile_gsweep.py 284160 10000 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gsweep_kernel -s 1 -c 1 -f -o gpurun_out/gsweep_pipe python tools/profile_gsweep.py 284160 10000 > gpurun_out/ncu_pipe.log 2>&1
tail -3 gpurun_out/ncu_pipe.log
