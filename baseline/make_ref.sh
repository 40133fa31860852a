#!/bin/bash
# Build-container helper: put a copy of the UNMODIFIED pure-Python reference under baseline/_ref/ (git-ignored, but NOT
# gpurun-ignored, so it travels to the GPU box with the snapshot).  The reference has no setup.py / pyproject, so there is
# nothing to pip-install: the tree is copied as it is (docs/ and .git left out).  Used ONLY by
#   * tests/test_gpu_planner_real.py   BASELINE config 4: the reference's own MotionPlannerGrad driven through dp.attach
#   * bench.py's `cpu_baseline.reference` leg and `--impl reference`: the reference's own compute_density on the host cores
# through oracle/ref_harness.py (DENSITY_PLANNER_REF points here).  Nothing in density_planner_b200/ imports it.
set -e
cd "$(dirname "$0")"
SRC=${1:-/root/reference}
[ -d "$SRC/systems" ] || { echo "no reference at $SRC"; exit 0; }
rm -rf _ref
mkdir -p _ref
tar -C "$SRC" --exclude=.git --exclude=docs -cf - . | tar -C _ref -xf -
echo "baseline/_ref: $(du -sh _ref | cut -f1) from $SRC"
