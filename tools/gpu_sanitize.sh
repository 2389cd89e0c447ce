#!/bin/bash
# compute-sanitizer on small renders through the CLI (one tool per gpurun call): usage tools/gpu_sanitize.sh memcheck|racecheck
set -u
TOOL=${1:-memcheck}
mkdir -p gpurun_out
OUT=gpurun_out/r02_sanitizer_$TOOL.txt
: > $OUT
CLI=./zig-weekend-raytracer_b200/weekend-raytracer
run() {
  echo "## compute-sanitizer --tool $TOOL $CLI $*" >> $OUT
  compute-sanitizer --tool $TOOL --error-exitcode 9 $CLI "$@" --image_out_path=gpurun_out/san.ppm >> $OUT 2>&1
  echo "## exit code $?" >> $OUT
}
run --image_width=100 --image_height=100 --samples_per_pixel=8 --ray_bounce_max_depth=10 --scene=emissive --writer=device
run --image_width=96 --image_height=54 --samples_per_pixel=4 --ray_bounce_max_depth=20 --scene=balls
run --image_width=96 --image_height=54 --samples_per_pixel=4 --ray_bounce_max_depth=20 --scene=synthetic --synthetic_prims=20000
rm -f gpurun_out/san.ppm
grep -E "^## |ERROR SUMMARY|RACECHECK SUMMARY" $OUT
