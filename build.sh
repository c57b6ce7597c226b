#!/bin/bash
# Build libtvts_b200.so (sm_100a) in-tree.  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
mkdir -p tvts_b200/lib build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets"
objs=""
pids=""
for f in tvts_b200/csrc/*.cu; do
  o=build/$(basename ${f%.cu}).o
  objs="$objs $o"
  if [ ! -f $o ] || [ $f -nt $o ] || [ tvts_b200/csrc/common.cuh -nt $o ] || [ tvts_b200/csrc/attention_common.cuh -nt $o ] || [ include/tvts_b200.h -nt $o ]; then
    fm="--use_fast_math"
    case $f in *loss.cu|*layernorm.cu|*optim.cu) fm="";; esac   # exact expf/logf/div where parity is tight
    $NVCC $FLAGS $fm ${PTXAS_V:+-Xptxas -v} -c $f -o $o &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -shared -o tvts_b200/lib/libtvts_b200.so $objs -Xcompiler -fPIC
echo "built tvts_b200/lib/libtvts_b200.so"
