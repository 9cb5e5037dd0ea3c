export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "stem" > gpurun_out/t40_stem.log 2>&1; echo "stem tests rc=$?"; tail -3 gpurun_out/t40_stem.log
timeout 600 python -m pytest tests/test_gpu_eco.py -x -q -m gpu > gpurun_out/t40_eco.log 2>&1; echo "eco tests rc=$?"; tail -3 gpurun_out/t40_eco.log
timeout 300 python tools/ab_bench.py --batch 32 --iters 3 base= > gpurun_out/ab40.log 2>&1; echo "ab rc=$?"; grep -E "^op|conv1|TOTAL|videos" gpurun_out/ab40.log | cut -c1-60
