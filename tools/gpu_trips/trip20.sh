export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 400 python tools/ab_bench.py --batch 32 --iters 3 base= print=debug_flags:16 notma=debug_flags:8 nomma=debug_flags:4 noldtm=debug_flags:2 nostore=debug_flags:1 alloff=debug_flags:15 > gpurun_out/ab20.log 2>&1; echo "ab rc=$?"; grep -m 8 "stem_rows cta0" gpurun_out/ab20.log; grep -A3 "^op " gpurun_out/ab20.log; tail -3 gpurun_out/ab20.log
bash tools/run_gpu_suite.sh "ops" 2>&1 | grep -E "rc=|passed|failed|error" | head -30
timeout 600 python -m pytest tests/test_gpu_eco.py -m gpu -q -k "test_eco_lite_n4_every_blob or test_eco_full" --timeout 300 > gpurun_out/t20_eco.log 2>&1; echo "eco rc=$?"; tail -5 gpurun_out/t20_eco.log
