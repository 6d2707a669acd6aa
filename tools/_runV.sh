L=online_convex_optimization_b200/csrc/liboco_b200.so
cp $L /tmp/lib_orig.so
for v in orig mb6 mb4 nf3; do
  if [ "$v" = orig ]; then cp /tmp/lib_orig.so $L; else cp online_convex_optimization_b200/csrc/_variants/lib_$v.so $L; fi
  echo "== $v"
  if [ "$v" != nf3 ]; then python tools/try_gsweep.py time 1000000 2>&1 | grep "mode=linear"; fi
  if [ "$v" = nf3 ] || [ "$v" = orig ]; then python tools/try_c3.py 1000000 2>&1 | grep "legacy=0"; fi
done
cp /tmp/lib_orig.so $L
