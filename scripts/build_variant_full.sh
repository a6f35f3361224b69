# full A/B build of the library with extra -D flags:  build_variant_full.sh <tag> -DMXB_BANK_BLOCK=256 ...
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
O=maximilian_b200/build/exp_$tag; X=maximilian_b200/lib_exp; mkdir -p $O $X
FL="-O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -gencode arch=compute_100a,code=sm_100a"
for f in maximilian_b200/csrc/*.cu; do
  nvcc $FL "$@" -c $f -o $O/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -o $X/libmaxib200_$tag.so $O/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
ls -la $X/libmaxib200_$tag.so
