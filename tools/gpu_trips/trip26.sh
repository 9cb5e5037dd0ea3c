export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t26_all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 gpurun_out/t26_all_gpu.log
timeout 400 python tools/ab_bench.py --batch 32 --iters 3 base= > gpurun_out/ab26.log 2>&1; echo "ab rc=$?"; grep -A40 "^op " gpurun_out/ab26.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench26.log 2>&1; echo "bench rc=$?"; tail -c 2900 gpurun_out/bench26.log
