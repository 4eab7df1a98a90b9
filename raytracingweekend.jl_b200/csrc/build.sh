#!/usr/bin/env bash
# Builds librtw_b200.so (the C-ABI library) for sm_100a, in-tree, next to the sources.
#   -fmad=false : nothing is contracted unless written as fmaf()/fma() (FP contract)
#   -lineinfo   : ncu source view maps SASS to these files
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false
       -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v)
SRCS=(rtw_kernels rtw_fused2 rtw_f64 rtw_capi rtw_image rtw_wavefront rtw_small rtw_scenegen)
# RTW_BUILD_VARIANTS=1: also build the kernel families kept only as measured comparisons (the first fused kernel with
# its rays-per-lane / sweep variants, the CTA wavefront, 4 cooperating lanes, per-slot candidate walks); their tests are
# marked `variants`.  The default library ships the default kernel, the split wavefront, the grid mode and Float64.
OUT=librtw_b200.so
OBJDIR=.
if [ "${RTW_BUILD_VARIANTS:-0}" = "1" ]; then
    FLAGS+=(-DRTW_BUILD_VARIANTS)
    SRCS+=(rtw_cta_wavefront)
    OUT=librtw_b200_variants.so   # select it with RTW_B200_LIB=<path>
    OBJDIR=variants_obj
    mkdir -p "$OBJDIR"
fi
pids=()
for s in "${SRCS[@]}"; do
    "$NVCC" "${FLAGS[@]}" -c "$s.cu" -o "$OBJDIR/$s.o" > "$s.log" 2>&1 &
    pids+=($!)
done
rc=0
for i in "${!pids[@]}"; do
    if ! wait "${pids[$i]}"; then rc=1; echo "== ${SRCS[$i]}.cu failed:"; cat "${SRCS[$i]}.log"; fi
done
[ $rc -eq 0 ] || exit 1
cat ./*.log | grep -E "spill|registers" | sort | uniq -c | sort -rn | head -5 || true
rm -f ./*.log
OBJS=()
for s in "${SRCS[@]}"; do OBJS+=("$OBJDIR/$s.o"); done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a "${OBJS[@]}" -o "$OUT" -lpthread -ldl
echo "built $(pwd)/$OUT"
