export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "stem" > gpurun_out/t41_stem.log 2>&1; echo "stem tests rc=$?"; tail -3 gpurun_out/t41_stem.log
timeout 300 python tools/ab_bench.py --batch 32 --iters 3 base= gw5=stem_gather_warps:5 gw6=stem_gather_warps:6 print=debug_flags:16 > gpurun_out/ab41.log 2>&1; echo "ab rc=$?"; grep -m 4 "stem_rows cta0" gpurun_out/ab41.log; grep -E "^op|conv1|TOTAL|videos" gpurun_out/ab41.log | cut -c1-90
timeout 600 python -m pytest tests/test_gpu_eco.py -x -q -m gpu > gpurun_out/t41_eco.log 2>&1; echo "eco tests rc=$?"; tail -3 gpurun_out/t41_eco.log
