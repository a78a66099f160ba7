#!/bin/bash
# Measurements this round ended before (no GPU time left): run on a B200 box from the repo root, results under gpurun_out/next/.
#   1. the one-lane-per-segment geometry (csrc/vit_kernel_l1.inc) against the product's 8-lane kernel by stream count and input
#      type (scripts/l1_bench.cu; one run so far: 16 s8 streams, 119.2 against 108.8 Gb/s, profiles/r2_l1_bench.txt);
#   2. ncu of the l1 kernel (issue-slot use, pipe utilisation, stall reasons: 3 warps per scheduler at 144 registers);
#   3. the 1024-stream job with larger waves (direct-store gather needs no per-wave step, so fewer, larger launches only
#      shorten the tails): --wave 16 / 32 against the default 8.
OUT=gpurun_out/next; mkdir -p $OUT
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo"
for S in 8 12 16 24 48; do
  for IN in IN_S8 IN_S4 IN_HARD; do
    $NV -DL1_STREAMS=$S -DL1_IN=vitk::$IN -o /tmp/l1_bench_${S}_$IN scripts/l1_bench.cu && /tmp/l1_bench_${S}_$IN > $OUT/l1_${S}_$IN.txt 2>&1
  done
done
$NV -DL1_STREAMS=16 -o /tmp/l1_bench scripts/l1_bench.cu
ncu --set full --clock-control none --import-source on -k regex:vit_decode_kernel_l1 -c 1 -o $OUT/l1_s8_16streams /tmp/l1_bench > $OUT/ncu_l1.log 2>&1
VIT_TEST_L1=1 python -m pytest tests/test_gpu_l1_geometry.py -q -m gpu > $OUT/pytest_l1.txt 2>&1    # L1 through the library
for W in 16 32; do
  VIT_GEOMETRY=l1 python bench.py --workload config5 --streams 128 --wave $W --batch 64 --gather none > $OUT/c5_128streams_l1_wave$W.json 2> $OUT/c5_l1_wave$W.err
done
for W in 8 16 32; do
  python bench.py --workload config5 --streams 128 --wave $W --batch 64 > $OUT/c5_128streams_wave$W.json 2> $OUT/c5_wave$W.err
done
ls $OUT
