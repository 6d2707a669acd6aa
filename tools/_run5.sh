L=online_convex_optimization_b200/csrc/liboco_b200.so
cp $L /tmp/orig.so
python tools/try_gsweep.py parity time 473600 1000000 2>&1 | tail -5
cp online_convex_optimization_b200/csrc/_variants/lib_mb4.so $L
python tools/try_gsweep.py time 378880 1000000 2>&1 | tail -3
cp /tmp/orig.so $L
