#!/bin/sh
# TEST INFRASTRUCTURE.  Builds oracle/_ref/libref_gnss.so from the reference's OWN sources where
# they lie under /root/reference (nothing is copied): RVI/gnss/src/lambda.cpp and
# RVI/gnss/src/common_function.cpp are plain C-style code; their shared header pulls in Eigen,
# marginalization_factor.h and ceres/problem.h only for type names, which oracle/ref_stubs/
# satisfies.  The rest of the reference (Ceres, factors, estimator) needs Eigen3/ROS/OpenCV and is
# NOT buildable in this image.  Outputs go to oracle/_ref/ only (git-ignored, travels with gpurun).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference/rtk_visual_inertial_src/rtk_visual_inertial/src/gnss
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$HERE/_ref"
g++ -O2 -fPIC -shared -ffp-contract=off -I"$HERE/ref_stubs" -I"$REF/include" \
    "$REF/src/lambda.cpp" "$REF/src/common_function.cpp" "$HERE/ref_shim.cpp" \
    -o "$HERE/_ref/libref_gnss.so"
echo "built $HERE/_ref/libref_gnss.so"
