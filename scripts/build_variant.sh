#!/bin/bash
# build_variant.sh <name> <extra nvcc flags...>: an A/B build of the CUDA library under gpurun_variants/<name>/ (travels to the
# GPU box with the snapshot; select it with DLSC_B200_LIB=gpurun_variants/<name>/libdlsc_b200.so)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/gpurun_variants/$name
make -C $root/dlsc_gc_planner_b200/csrc OUT=$root/gpurun_variants/$name/libdlsc_b200.so OBJ=$root/gpurun_variants/$name/build EXTRA="$*" > /dev/null
echo built $name
