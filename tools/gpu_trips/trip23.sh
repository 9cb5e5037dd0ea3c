export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "test_inception" --timeout 300 -x > gpurun_out/t23_inc.log 2>&1; echo "inception rc=$?"; tail -30 gpurun_out/t23_inc.log
timeout 900 python -m pytest tests/test_gpu_eco.py -m gpu -q --timeout 600 > gpurun_out/t23_eco.log 2>&1; echo "eco rc=$?"; tail -12 gpurun_out/t23_eco.log
timeout 400 python tools/ab_bench.py --batch 32 --iters 3 base= nofuse=fuse_1x1:0 > gpurun_out/ab23.log 2>&1; echo "ab rc=$?"; grep -A48 "^op " gpurun_out/ab23.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench23.log 2>&1; echo "bench rc=$?"; tail -c 2600 gpurun_out/bench23.log
bash tools/run_gpu_suite.sh "ops" 2>&1 | grep -E "rc=|passed|failed|error" | head -30
