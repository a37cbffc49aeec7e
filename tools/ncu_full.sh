#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each) from the eager B=1 full-pipeline step.
# usage: tools/ncu_full.sh <tag>
TAG=${1:-r01}; OUT=gpurun_out; mkdir -p $OUT
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $OUT/${TAG}_full_$1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-trace --no-graph > $OUT/${TAG}_full_$1.log 2>&1
  echo "$1 rc=$?"
}
WHAT=${2:-all}
if [ "$WHAT" = all ] || [ "$WHAT" = conv ]; then
cap conv_persistent "conv_gemm_persistent_kernel" 1     # VAE encoder resnet conv 128->128 @512^2, 4 reference images
cap conv_onetile "^conv_gemm_kernel" 0
fi
if [ "$WHAT" = all ] || [ "$WHAT" = rest ]; then
cap attn "shared_attn_kernel" 0
cap gn_apply "gn_apply_kernel" 2
fi
ls -la $OUT/${TAG}_full_*.ncu-rep
