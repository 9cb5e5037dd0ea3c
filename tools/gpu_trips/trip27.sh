export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "pool or inception" > gpurun_out/t27_pool.log 2>&1; echo "pool tests rc=$?"; tail -3 gpurun_out/t27_pool.log
timeout 400 python tools/ab_bench.py --batch 32 --iters 3 base= > gpurun_out/ab27.log 2>&1; echo "ab rc=$?"; grep -E "^op|pool|TOTAL|videos" gpurun_out/ab27.log
timeout 400 python tools/ab_bench.py --batch 32 --iters 3 --model full base= > gpurun_out/ab27_full.log 2>&1; echo "ab full rc=$?"; grep -A80 "^op " gpurun_out/ab27_full.log | cut -c1-60
