mkdir -p gpurun_out
python tools/sanitize_smoke.py 2>&1 | tail -2
python tools/profile_fill.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fill_warp_kernel -s 1 -c 1 -f -o gpurun_out/r02_fill_warp python tools/profile_fill.py > gpurun_out/ncu_fill.log 2>&1
tail -1 gpurun_out/ncu_fill.log
