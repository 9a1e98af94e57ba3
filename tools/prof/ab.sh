# same-box A/B of two builds of the library: tools/prof/ab.sh (expects csrc/libgmr_old.so next to libgmr_b200.so)
for lib in libgmr_old.so libgmr_b200.so libgmr_old.so libgmr_b200.so; do
  echo "== $lib"
  GMR_B200_LIB=$PWD/general_motion_retargeting_b200/csrc/$lib python tools/prof/latency_probe.py 2>/dev/null | python tools/prof/fmt_probe.py | grep -E "(lone_warp|balanced_all|bench)"
done
