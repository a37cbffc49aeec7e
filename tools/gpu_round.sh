#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list (+ optional full capture). Logs -> gpurun_out/.
# usage: tools/gpu_round.sh <tag> [tests|bench|ncu|caps|full ...]   (default: tests bench ncu)
set -u
TAG=${1:-r01}; shift || true
WHAT=${*:-tests bench ncu}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/${TAG}_smi.csv 2>&1
for w in $WHAT; do
case $w in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
  tail -5 $OUT/${TAG}_pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_smoke.log ;;
bench)
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/${TAG}_clocks.csv 2>/dev/null &
  SMI=$!
  timeout 1200 python bench.py --steps 20 --warmup 5 --trace-out $OUT/${TAG}_kernels.json > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
  kill $SMI 2>/dev/null
  head -c 3000 $OUT/${TAG}_bench.json; echo; tail -3 $OUT/${TAG}_bench.err ;;
bench8)
  timeout 600 python bench.py --steps 10 --warmup 3 --batch 8 --no-cpu-baseline --no-extras --trace-out $OUT/${TAG}_kernels_b8.json > $OUT/${TAG}_bench_b8.json 2> $OUT/${TAG}_bench_b8.err; echo "bench b8 rc=$?"
  head -c 1500 $OUT/${TAG}_bench_b8.json; echo; tail -3 $OUT/${TAG}_bench_b8.err ;;
ab)
  # same-box A/B of an environment switch: AB_ENV="IR_GN_SINGLE_LAUNCH=0" tools/gpu_round.sh <tag> ab
  for B in 1 8; do for arm in "" "${AB_ENV:-IR_GN_SINGLE_LAUNCH=0}"; do
    nm=$(echo "${arm:-default}" | tr '/' '_' | sed 's/.*instantrestore_b200_//')
    env $arm timeout 600 python bench.py --steps 20 --warmup 3 --batch $B --no-cpu-baseline --no-extras --no-trace > $OUT/${TAG}_ab_b${B}_$nm.json 2>> $OUT/${TAG}_ab.err
    python -c "import json,sys; d=json.load(open('$OUT/${TAG}_ab_b${B}_$nm.json')); print('AB B=$B', '$nm', round(d['value'],2), 'images/s', round(d['ms_per_step'],3), 'ms', d['gpu_launches_per_step'], 'launches', d['clocks'])"
  done; done ;;
streams)
  for S in 2 3 4 5 6; do
    timeout 300 python bench.py --steps 30 --warmup 3 --streams $S --no-cpu-baseline --no-extras --no-trace > $OUT/${TAG}_streams_$S.json 2>> $OUT/${TAG}_streams.err
    python -c "import json; d=json.load(open('$OUT/${TAG}_streams_$S.json')); print('STREAMS $S', round(d['value'],2), 'images/s e2e', round(d['e2e']['value'],2), d['clocks'])"
  done ;;
ncu)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-trace --no-graph > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
  wc -l $OUT/${TAG}_launches.csv ;;
caps|capsattn|capsnew|capsplain|capshalo)
  # ncu --set full of the dominant kernels on micro-drivers; summarised on the box (reports are ~10 MB each and
  # gpurun_out is capped at 64 MiB), only the text summaries and ncu_traffic.json travel back
  cap() { n=$1; k=$2; key=$3; shift 3
    timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/${TAG}_$n python "$@" > $OUT/${TAG}_cap_$n.log 2>&1
    echo "cap $n rc=$?"
    ncu -i /tmp/${TAG}_$n.ncu-rep --page raw --csv > $OUT/${TAG}_raw_$n.csv 2>/dev/null
    python tools/ncu_summary.py /tmp/${TAG}_$n.ncu-rep profiles/${TAG}_ncu_full_$n.txt "$key" > /dev/null && cp profiles/${TAG}_ncu_full_$n.txt profiles/ncu_traffic.json $OUT/
    rm -f /tmp/${TAG}_$n.ncu-rep; }
  if [ $w = capsnew ]; then
  # round-2 additions: the folded upsamplers of the VAE decoder and the TMA-store epilogue on the patch conv_in
  cap conv_up2x_512 conv "ir_conv_gemm:m262144_k2048_n512_ks3s1_up2x" tools/gemm_one.py up 4 128 512 512
  cap conv_up2x_256 conv "ir_conv_gemm:m1048576_k1024_n256_ks3s1_up2x" tools/gemm_one.py up 4 256 256 256
  cap conv_in_tma_store conv "ir_conv_gemm:m1048576_k64_n128_ks1s1" tools/gemm_one.py lin_ts 4 1048576 64 128 0
  cap conv_in_row_store conv "ir_conv_gemm:m1048576_k64_n128_ks1s1_rowstores" tools/gemm_one.py lin_ts 4 1048576 64 128 1
  elif [ $w = capshalo ]; then
  # 128-channel layer (128->128 @512^2, 4 images, residual): the 128-wide CTA pair with the TMA-in / TMA-out epilogue, and the single-CTA kernel
  cap conv_halo_pair128 conv3_halo "ir_conv_gemm:m1048576_k1152_n128_ks3s1" tools/gemm_one.py conv3 4 512 128 128
  cap conv_halo_128_single conv3_halo "ir_conv_gemm:m1048576_k1152_n128_ks3s1_single" tools/gemm_one.py conv3 4 512 128 128 1
  elif [ $w = capsplain ]; then
  # the plain variant on three rotating score buffers, next to the (unchanged) shared-image variant
  cap attn_b32 shared_attn "ir_shared_attn_fwd:b32_h5_sq4096_skv4096" tools/attn_one.py 32 5 4096 1 0 0
  cap attn_b4 shared_attn "ir_shared_attn_fwd:b4_h5_sq4096_skv4096" tools/attn_one.py 4 5 4096 1 0 0
  cap attn_shared_b1 shared_attn "ir_shared_attn_fwd:b1_h5_sq4096_skv16384_adain" tools/attn_one.py 1 5 4096 0 4 1
  elif [ $w = capsattn ]; then
  cap attn_shared_b1 shared_attn "ir_shared_attn_fwd:b1_h5_sq4096_skv16384_adain" tools/attn_one.py 1 5 4096 0 4 1
  cap attn_shared_b8 shared_attn "ir_shared_attn_fwd:b8_h5_sq4096_skv16384_adain" tools/attn_one.py 8 5 4096 0 4 1
  cap gn_apply_512 gn_apply "ir_groupnorm:b4_hw262144_c128" tools/norm_one.py gn 4 262144 128
  else
  cap conv_halo_pair_512 conv "ir_conv_gemm:m65536_k4608_n512_ks3s1" tools/gemm_one.py conv3 4 128 512 512
  cap conv_halo_128 conv "ir_conv_gemm:m1048576_k1152_n128_ks3s1" tools/gemm_one.py conv3 4 512 128 128
  cap conv_pair160_320 conv "ir_conv_gemm:m131072_k2880_n320_ks3s1" tools/gemm_one.py conv3 32 64 320 320
  cap conv_onetile_1280 conv_gemm_kernel "ir_conv_gemm:m1024_k1280_n1280_ks1s1" tools/gemm_one.py lin 1 1024 1280 1280
  cap attn_b32 shared_attn "ir_shared_attn_fwd:b32_h5_sq4096_skv4096" tools/attn_one.py 32 5 4096 1 0 0
  cap attn_b4 shared_attn "ir_shared_attn_fwd:b4_h5_sq4096_skv4096" tools/attn_one.py 4 5 4096 1 0 0
  # the shared-image variant (<ADAIN>: reference chunks, AdaIN affine, split-KV at one identity per step) and its B = 8 launch
  cap attn_shared_b1 shared_attn "ir_shared_attn_fwd:b1_h5_sq4096_skv16384_adain" tools/attn_one.py 1 5 4096 0 4 1
  cap attn_shared_b8 shared_attn "ir_shared_attn_fwd:b8_h5_sq4096_skv16384_adain" tools/attn_one.py 8 5 4096 0 4 1
  cap gn_apply_512 gn_apply "ir_groupnorm:b4_hw262144_c128" tools/norm_one.py gn 4 262144 128
  cap gn_partial_512 gn_partial "ir_groupnorm_partial:b4_hw262144_c128" tools/norm_one.py gn 4 262144 128
  cap gn_single_320 gn_fused "ir_groupnorm:b4_hw4096_c320_1launch" tools/norm_one.py gn 4 4096 320 2
  cap layernorm_320 layernorm "ir_layernorm:r16384_c320" tools/norm_one.py ln 16384 320
  fi ;;
full)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-shared_attn} -s ${NCU_SKIP:-20} -c ${NCU_COUNT:-3} \
     -f -o $OUT/${TAG}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-trace --no-graph > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
  ls -la $OUT/${TAG}_prof* ;;
esac
done
