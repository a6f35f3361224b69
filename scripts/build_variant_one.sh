# A/B build of the library where only some sources are recompiled with extra -D flags (the other objects come from the regular build):
#   build_variant_one.sh <tag> "<src1.cu src2.cu ...>" -DMACRO=...
set -e
cd "$(dirname "$0")/.."
tag=$1; srcs=$2; shift 2
O=maximilian_b200/build/exp_$tag; X=maximilian_b200/lib_exp; mkdir -p $O $X
FL="-O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -gencode arch=compute_100a,code=sm_100a"
objs=""
for f in maximilian_b200/csrc/*.cu; do
  b=$(basename ${f%.cu})
  if echo " $srcs " | grep -q " $b.cu "; then nvcc $FL "$@" -c $f -o $O/$b.o & objs="$objs $O/$b.o"; else objs="$objs maximilian_b200/build/$b.o"; fi
done
wait
nvcc -shared -o $X/libmaxib200_$tag.so $objs -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
ls -la $X/libmaxib200_$tag.so
