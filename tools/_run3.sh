L=online_convex_optimization_b200/csrc/liboco_b200.so
cp $L /tmp/orig.so
for v in orig mb4 mb5 mb8; do
  if [ $v = orig ]; then cp /tmp/orig.so $L; B="113664 227328"; else cp online_convex_optimization_b200/csrc/_variants/lib_$v.so $L; fi
  case $v in mb4) B=75776;; mb5) B=94720;; mb8) B=151552;; esac
  echo "== $v"; timeout 300 python tools/try_gsweep_occ.py $B 2>&1 | grep debug_phase
done
cp /tmp/orig.so $L
