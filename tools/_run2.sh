set -x
mkdir -p gpurun_out
python tools/profile_gsweep.py 113664 10000 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gsweep_kernel -s 1 -c 1 -f -o gpurun_out/gsweep_f1 python tools/profile_gsweep.py 113664 10000 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gsweep_kernel -s 1 -c 1 -f -o gpurun_out/gsweep_f0 python tools/profile_gsweep.py 113664 10000 0 > gpurun_out/ncu0.log 2>&1
ls -la gpurun_out
