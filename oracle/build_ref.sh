#!/bin/sh
# TEST INFRASTRUCTURE.  Builds oracle/_ref/{libref_gnss.so, libswgn_refdemo.so, libref_estimator.so} from the reference's OWN
# sources where they lie under /root/reference (nothing is copied):
#   RVI/gnss/src/lambda.cpp, RVI/gnss/src/common_function.cpp         plain C-style code
#   RVI/factor/gnss_factor.cpp, projection_factor.cpp, imu_factor.cpp, integration_base.cpp,
#   pose_local_parameterization.cpp                                   the factor classes of the hot path
#   RVI/factor/initial_factor.cpp, pose0_factor.cpp                   initialisation factors (host-evaluated through the shim)
#   RVI/factor/marginalization_factor.cpp                             MarginalizationInfo::marginalize / MarginalizationFactor
#   RVI/factor/gnss_imu_factor.cpp                                    IMUGNSSBase / IMUGNSSFactor (the hidden GNSS-frame chain), through
#                                                                     this repository's include/ceres/{small_blas,invert_psd_matrix}.h
# Eigen, OpenCV and Ceres are not installed in this image: the factor sources are compiled against
# oracle/ref_stubs/ (a minimal eager stand-in for the part of Eigen's dense API they use, empty OpenCV
# headers, a type-name stub of marginalization_factor.h) and against this repository's own
# include/ceres/ headers (CostFunction / SizedCostFunction / LocalParameterization), which doubles as
# the build test of those headers against unmodified reference code.  NOT buildable here, and not
# attempted: the estimator (ROS, OpenCV) and Ceres itself.
# Outputs go to oracle/_ref/ only (git-ignored, travels with gpurun).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC=/root/reference/rtk_visual_inertial_src/rtk_visual_inertial/src
REF=$SRC/gnss
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$HERE/_ref"
# (-include: the real <ceres/ceres.h> drags Eigen and <numeric> in, which marginalization_factor.h relies on)
g++ -O2 -fPIC -shared -ffp-contract=off -std=c++14 -include numeric -include eigen3/Eigen/Dense \
    -I"$HERE/ref_stubs" -I"$HERE/../include" -I"$REF/include" -I"$SRC" \
    "$REF/src/lambda.cpp" "$REF/src/common_function.cpp" \
    "$SRC/factor/gnss_factor.cpp" "$SRC/factor/projection_factor.cpp" "$SRC/factor/imu_factor.cpp" \
    "$SRC/factor/integration_base.cpp" "$SRC/factor/pose_local_parameterization.cpp" \
    "$SRC/factor/initial_factor.cpp" "$SRC/factor/marginalization_factor.cpp" "$SRC/factor/gnss_imu_factor.cpp" \
    "$HERE/ref_shim.cpp" "$HERE/ref_marg_shim.cpp" "$HERE/ref_globals.cpp" "$HERE/ref_problem_stubs.cpp" -lpthread -o "$HERE/_ref/libref_gnss.so"
echo "built $HERE/_ref/libref_gnss.so"
# The drop-in demonstration of the ceres:: shim on the reference's own factor classes (shim/ceres_shim_refdemo.cpp):
# needs the product library (libswgn.so) and the generator, built by __graft_entry__.build() before this script runs.
PKG="$HERE/../rtk-visual-inertial-navigation_b200"
if [ -f "$PKG/libswgn.so" ] && [ -f "$PKG/libswgn_synth.so" ]; then
  g++ -O2 -fPIC -shared -ffp-contract=off -std=c++17 -include numeric -include eigen3/Eigen/Dense -I"$HERE/ref_stubs" -I"$HERE/../include" -I"$REF/include" -I"$SRC" -I"$PKG/shim" \
      "$SRC/factor/gnss_factor.cpp" "$SRC/factor/projection_factor.cpp" "$SRC/factor/imu_factor.cpp" \
      "$SRC/factor/integration_base.cpp" "$SRC/factor/pose_local_parameterization.cpp" "$REF/src/common_function.cpp" \
      "$SRC/factor/initial_factor.cpp" "$SRC/factor/pose0_factor.cpp" "$SRC/factor/marginalization_factor.cpp" \
      "$SRC/factor/gnss_imu_factor.cpp" \
      "$PKG/shim/ceres_shim.cpp" "$PKG/shim/ceres_shim_refdemo.cpp" "$PKG/shim/gnss_refdemo.cpp" "$HERE/ref_globals.cpp" \
      -o "$HERE/_ref/libswgn_refdemo.so" -L"$PKG" -lswgn -lswgn_synth -Wl,-rpath,'$ORIGIN/../../rtk-visual-inertial-navigation_b200'
  echo "built $HERE/_ref/libswgn_refdemo.so"
  # The reference's estimator code for the per-epoch GNSS path (swf_gnss.cpp, swf_core.cpp, swf_lambda.cpp), unmodified, on the
  # ceres:: shim: oracle/ref_estimator_shim.cpp supplies what lives in translation units that need ROS / OpenCV.
  g++ -O2 -fPIC -shared -ffp-contract=off -std=c++17 -include numeric -include eigen3/Eigen/Dense -I"$HERE/ref_stubs" -I"$HERE/../include" \
      -I"$REF/include" -I"$SRC" -I"$SRC/factor" -I"$PKG/shim" \
      "$SRC/swf/swf_gnss.cpp" "$SRC/swf/swf_core.cpp" "$SRC/swf/swf_lambda.cpp" \
      "$SRC/factor/gnss_factor.cpp" "$SRC/factor/projection_factor.cpp" "$SRC/factor/imu_factor.cpp" \
      "$SRC/factor/integration_base.cpp" "$SRC/factor/pose_local_parameterization.cpp" "$SRC/factor/initial_factor.cpp" \
      "$SRC/factor/pose0_factor.cpp" "$SRC/factor/mag_factor.cpp" "$SRC/factor/marginalization_factor.cpp" \
      "$SRC/factor/gnss_imu_factor.cpp" "$REF/src/common_function.cpp" "$REF/src/lambda.cpp" \
      "$PKG/shim/ceres_shim.cpp" "$PKG/shim/ceres_shim_refdemo.cpp" "$HERE/ref_estimator_shim.cpp" "$HERE/ref_globals.cpp" \
      -o "$HERE/_ref/libref_estimator.so" -L"$PKG" -lswgn -lswgn_synth -lpthread -Wl,-rpath,'$ORIGIN/../../rtk-visual-inertial-navigation_b200' -Wl,-z,defs
  echo "built $HERE/_ref/libref_estimator.so"
fi
