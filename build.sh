#!/bin/bash
# Build the CUDA library (sm_100a) in-tree.  Used by __graft_entry__.build().
#   ./build.sh                     -> tvts_b200/lib/libtvts_b200.so        (bf16 operands, the default)
#   TVTS_OPERAND=fp16 ./build.sh   -> tvts_b200/lib/libtvts_b200_fp16.so   (IEEE-half operands: -DTVTS_OPERAND_FP16)
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets"
BUILD=build
LIB=tvts_b200/lib/libtvts_b200.so
if [ "${TVTS_OPERAND:-bf16}" = "fp16" ]; then
  FLAGS="$FLAGS -DTVTS_OPERAND_FP16"
  BUILD=build_fp16
  LIB=tvts_b200/lib/libtvts_b200_fp16.so
fi
mkdir -p tvts_b200/lib $BUILD
objs=""
pids=""
for f in tvts_b200/csrc/*.cu; do
  o=$BUILD/$(basename ${f%.cu}).o
  objs="$objs $o"
  if [ ! -f $o ] || [ $f -nt $o ] || [ tvts_b200/csrc/common.cuh -nt $o ] || [ tvts_b200/csrc/attention_common.cuh -nt $o ] || [ include/tvts_b200.h -nt $o ]; then
    fm="--use_fast_math"
    case $f in *loss.cu|*loss_fused.cu|*layernorm.cu|*optim.cu|*input_stage.cu) fm="";; esac   # exact expf/logf/div where parity is tight
    $NVCC $FLAGS $fm ${PTXAS_V:+-Xptxas -v} -c $f -o $o &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -shared -o $LIB $objs -Xcompiler -fPIC -ldl
echo "built $LIB"
