# Builds a variant of the library with extra nvcc flags into is-dqn_b200/lib/libisdqn_b200_$1.so (objects in build_$1/):
#   scripts/build_variant.sh pdl -DISDQN_PDL_LATE=1   then   ISDQN_LIB=$PWD/is-dqn_b200/lib/libisdqn_b200_pdl.so ISDQN_PDL=1 python bench.py
set -e
NAME="$1"; shift
mkdir -p build_$NAME
for f in is-dqn_b200/csrc/*.cu; do
  o=build_$NAME/$(basename ${f%.cu}).o
  nvcc "$@" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -c $f -o $o &
done
wait
nvcc -shared -o is-dqn_b200/lib/libisdqn_b200_$NAME.so build_$NAME/*.o -ldl -Xlinker --version-script=is-dqn_b200/csrc/exports.map
ls -la is-dqn_b200/lib/
