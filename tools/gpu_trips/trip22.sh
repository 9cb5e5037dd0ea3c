export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
bash tools/run_gpu_suite.sh "ops" 2>&1 | grep -E "rc=|passed|failed|error" | head -30
grep -E "Error|assert|FAILED" gpurun_out/suite_ops_test_inception.log gpurun_out/suite_ops_test_stem.log | head -20
timeout 900 python -m pytest tests/test_gpu_eco.py -m gpu -q --timeout 600 > gpurun_out/t22_eco.log 2>&1; echo "eco rc=$?"; tail -12 gpurun_out/t22_eco.log
timeout 400 python tools/ab_bench.py --batch 32 --iters 3 base= print=debug_flags:16 nocommute=pool_commute:0 > gpurun_out/ab22.log 2>&1; echo "ab rc=$?"; grep -m 4 "stem_rows cta0" gpurun_out/ab22.log; grep -A48 "^op " gpurun_out/ab22.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench22.log 2>&1; echo "bench rc=$?"; tail -c 2600 gpurun_out/bench22.log
