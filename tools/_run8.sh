L=online_convex_optimization_b200/csrc/liboco_b200.so
cp $L /tmp/orig.so
python tools/try_c3.py parity 1000000 2>&1 | tail -8
cp online_convex_optimization_b200/csrc/_variants/lib_nf5.so $L
python tools/try_c3.py 1000000 2>&1 | tail -2
cp /tmp/orig.so $L
