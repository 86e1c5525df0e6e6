#!/bin/bash
# A/B of two builds of libtik.so on the SAME box: per-kernel times of one configs[2] step (TIK_PLAN_TRACE), alternating
# the product library and a variant (python -m ...build with TIK_BUILD_VARIANT=<macro>), three rounds each.
#   tools/ab_trace.sh temporal_inverse_kinematics_b200/build/libtik_<variant>.so
VARIANT=$1
for r in 1 2 3; do
  for which in product variant; do
    if [ $which = variant ]; then export TIK_LIB_PATH=$PWD/$VARIANT; else unset TIK_LIB_PATH; fi
    TIK_PLAN_TRACE=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-hbm --no-c3 2>&1 >/dev/null | grep "tik trace" | tail -17 |
      awk -v w=$which -v r=$r '{printf "%s ", $(NF-1); s+=$(NF-1)} END {printf " | sum %.1f us  %s round %s\n", s, w, r}'
  done
done
