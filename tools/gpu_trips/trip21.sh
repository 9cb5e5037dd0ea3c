export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 120 ./build/probe_commit > gpurun_out/probe_commit2.log 2>&1; echo "probe rc=$?"; grep -A40 "mode 1" gpurun_out/probe_commit2.log | head -40
bash tools/run_gpu_suite.sh "ops" 2>&1 | grep -E "rc=|passed|failed|error" | head -30
timeout 900 python -m pytest tests/test_gpu_eco.py -m gpu -q --timeout 600 > gpurun_out/t21_eco.log 2>&1; echo "eco rc=$?"; tail -5 gpurun_out/t21_eco.log
timeout 400 python tools/ab_bench.py --batch 32 --iters 3 base= print=debug_flags:16 nopersist=persistent:0 allpersist=persistent:2 > gpurun_out/ab21.log 2>&1; echo "ab rc=$?"; grep -m 4 "stem_rows cta0" gpurun_out/ab21.log; grep -A45 "^op " gpurun_out/ab21.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench21.log 2>&1; echo "bench rc=$?"; tail -c 2600 gpurun_out/bench21.log
