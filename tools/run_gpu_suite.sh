#!/bin/bash
# Runs on the GPU box (under gpurun): every stage in its own process with its own timeout so one
# trapped kernel cannot take the rest of the log with it.  Output: gpurun_out/suite_*.log
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests:$PYTHONPATH"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/suite_gpu.log 2>&1
nproc >> gpurun_out/suite_gpu.log; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/suite_gpu.log
STAGES="${1:-probe debug ops eco}"
for st in $STAGES; do
  case $st in
    probe) timeout 120 ./build/probe_im2col > gpurun_out/suite_probe.log 2>&1; echo "probe rc=$?" ;;
    debug) for m in 0 1; do for c in gemm c3x3 d3 stem; do
             timeout 180 python tools/debug_conv.py $m $c > gpurun_out/suite_debug_${c}_m${m}.log 2>&1; echo "debug $c m$m rc=$?"; done; done ;;
    ops)   for k in test_conv2d test_conv3d test_stem test_generic test_conv_without test_residual test_inception test_pooling test_pool_variants; do
             timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k $k --timeout 300 -x > gpurun_out/suite_ops_$k.log 2>&1; echo "ops $k rc=$?"; done ;;
    eco)   for k in test_eco_lite_n4_every_blob test_eco_lite_n4_fast test_eco_full_n4 test_eco_lite_n16 test_pipelined; do
             timeout 900 python -m pytest tests/test_gpu_eco.py -m gpu -q -k $k --timeout 600 > gpurun_out/suite_eco_$k.log 2>&1; echo "eco $k rc=$?"; done ;;
    bench) timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/suite_bench.log 2>&1; echo "bench rc=$?" ;;
  esac
done
tail -n 6 gpurun_out/suite_*.log
