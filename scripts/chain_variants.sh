# diagnostics: chaining throughput for kernel variants built as csrc/libvariant_<minblocks>.so
for mb in 6 8 12; do
  cp gmap_2024_b200/csrc/libvariant_$mb.so gmap_2024_b200/csrc/libgmapdp_b200.so
  echo "== min blocks $mb"
  timeout 200 python scripts/chain_speed.py 256 128 2>&1 | grep "device:"
done
