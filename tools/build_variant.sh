#!/bin/bash
# Builds a variant of liberl_gp_b200.so for kernel experiments: the translation units named after the flags are recompiled
# with the extra -D flags, every other object comes from the regular build (run `make -C erl_gaussian_process_b200/csrc` first).
#   tools/build_variant.sh <name> "<-D flags>" <tu> [<tu> ...]     ->  erl_gaussian_process_b200/lib/liberl_gp_b200_<name>.so
# Use it with ERL_GP_B200_LIB=erl_gaussian_process_b200/lib/liberl_gp_b200_<name>.so.
set -e
cd "$(dirname "$0")/../erl_gaussian_process_b200/csrc"
name=$1; flags=$2; shift 2
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
OBJDIR=../../build/csrc
VDIR=../../build/variant_$name
mkdir -p $VDIR
objs=""
for f in *.cu; do
  tu=${f%.cu}
  if [[ " $* " == *" $tu "* ]]; then
    $NVCC $ARCH -O3 -std=c++17 -lineinfo -ccbin $(command -v g++) -Xcompiler -fPIC,-Wall,-fvisibility=hidden -I../../include -I. --expt-relaxed-constexpr -Xptxas -v -Xfatbin=-compress-all $flags -c $f -o $VDIR/$tu.o 2> $VDIR/$tu.ptxas.log || { cat $VDIR/$tu.ptxas.log; exit 1; }
    objs="$objs $VDIR/$tu.o"
  else
    objs="$objs $OBJDIR/$tu.o"
  fi
done
$NVCC $ARCH -shared -ccbin $(command -v g++) -o ../lib/liberl_gp_b200_$name.so $objs -cudart static
echo "built erl_gaussian_process_b200/lib/liberl_gp_b200_$name.so"
