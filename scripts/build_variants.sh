#!/bin/bash
# A/B builds of libb200sim.so with build-time knobs of the traversal (development aid; see csrc/nbody.cuh):
#   bash scripts/build_variants.sh name1 "-DB200_TRAV_CAP=448" name2 "-DB200_TRAV64_CTAS=4" ...
# -> scripts/bin/lib_<name>.so (git-ignored, ships with gpurun); load one with B200SIM_LIB=scripts/bin/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
src=3d-spatial-sim-for-boid-and-nbody_b200/csrc
out=scripts/bin
mkdir -p $out/obj
flags="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden"
for f in multi boids generate capi; do
    if [ ! -f $out/obj/$f.o ] || [ $src/$f.cu -nt $out/obj/$f.o ] || [ include/b200sim.h -nt $out/obj/$f.o ]; then
        (cd $src && nvcc $flags -c $f.cu -o ../../$out/obj/$f.o) &
    fi
done
wait
while [ $# -ge 2 ]; do
    name=$1; defs=$2; shift 2
    (cd $src && nvcc $flags $defs -Xptxas -v -c nbody.cu -o ../../$out/obj/nbody_$name.o 2> ../../$out/obj/nbody_$name.ptxas \
        && nvcc -shared -o ../../$out/lib_$name.so ../../$out/obj/nbody_$name.o ../../$out/obj/multi.o ../../$out/obj/boids.o ../../$out/obj/generate.o ../../$out/obj/capi.o -ldl \
        && echo "built $out/lib_$name.so [$defs]: $(grep -A2 'traverse64c_kernelILb0ELb1' ../../$out/obj/nbody_$name.ptxas | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')") &
done
wait
