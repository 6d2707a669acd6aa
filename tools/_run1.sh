set -x
L=online_convex_optimization_b200/csrc/liboco_b200.so
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 600 python tools/try_gsweep.py parity time > gpurun_out/gsweep_inline.log 2>&1; echo rc=$?
cp $L /tmp/orig.so; cp online_convex_optimization_b200/csrc/_variants/lib_noinl.so $L
timeout 600 python tools/try_gsweep.py parity time > gpurun_out/gsweep_noinl.log 2>&1; echo rc=$?
cp /tmp/orig.so $L
tail -30 gpurun_out/gsweep_inline.log; tail -12 gpurun_out/gsweep_noinl.log
