#!/bin/bash
# compute-sanitizer over the kernel-level parity tests: memcheck on every kernel family at small
# shapes, racecheck (shared-memory hazards) and synccheck on the kernels that synchronise through shared memory /
# clusters without the async proxy (norms, AdaIN statistics, concat, layout kernels). Summaries -> gpurun_out/<tag>_san_*.txt
# usage: tools/sanitize.sh <tag> [new|last]
TAG=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
SMALL='groupnorm or layernorm or adain or concat or upsample or latent_in or softmax'
run() {  # name tool timeout pytest-args...
  n=$1; tool=$2; to=$3; shift 3
  timeout $to $SAN --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest "$@" -q -x -m gpu -p no:cacheprovider > $OUT/${TAG}_san_$n.log 2>&1
  rc=$?
  { echo "# compute-sanitizer --tool $tool python -m pytest $* -m gpu"; echo "exit code: $rc (99 = sanitizer errors, 124 = timeout)";
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard|error" $OUT/${TAG}_san_$n.log | sort | uniq -c | sort -rn | head -15; } > $OUT/${TAG}_san_$n.txt
  cat $OUT/${TAG}_san_$n.txt
}
if [ "${2:-all}" = last ]; then  # the kernels of the last session: three-buffer attention, 128-wide halo pair (TMA-in / TMA-out epilogue)
  run attn3_memcheck memcheck 600 tests/test_gpu_kernels.py -k "attention"
  run halo_pair128_memcheck memcheck 600 tests/test_gpu_kernels.py -k "halo or pass_a"
  exit 0
fi
if [ "${2:-all}" = new ]; then   # the kernels added last: folded upsamplers, TMA-store epilogue, one-launch FreeU, image patches
  run new_memcheck memcheck 1200 tests/test_gpu_kernels.py -k "upsample2x or tma_store or concat_freeu or image_patches"
  run new_racecheck racecheck 600 tests/test_gpu_kernels.py -k "concat_freeu or image_patches"
  run faceid_memcheck memcheck 900 tests/test_gpu_pipeline.py -k "faceid and eager"
  exit 0
fi
run norms_memcheck memcheck 900 tests/test_gpu_kernels.py -k "$SMALL"
# racecheck does not model tcgen05.alloc's shared-memory write (it reports it against the read that follows the barrier +
# tcgen05 fences), so the GEMM-launching tests are left out of this pass
run norms_racecheck racecheck 900 tests/test_gpu_kernels.py -k "($SMALL) and not gemm_epilogue and not from_epilogue"
run norms_synccheck synccheck 900 tests/test_gpu_kernels.py -k "groupnorm_single_launch"
run gemm_memcheck memcheck 1200 tests/test_gpu_kernels.py -k "test_linear or test_geglu or test_conv3x3 or pass_a"
run attn_memcheck memcheck 900 tests/test_gpu_kernels.py -k "shared_attention and not large"
run image_memcheck memcheck 900 tests/test_preprocess.py
run pipeline_memcheck memcheck 1500 tests/test_gpu_pipeline.py -k "tiny_pipeline or processors_drive"
