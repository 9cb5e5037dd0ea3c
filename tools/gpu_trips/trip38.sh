export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t38_all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/t38_all_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r01.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r01_default.log 2>&1; echo "bench rc=$?"; tail -c 2900 gpurun_out/bench_r01_default.log
timeout 300 python tools/ab_bench.py --batch 32 --iters 3 base= > gpurun_out/ab38.log 2>&1; echo "ab rc=$?"; grep -E "3b_1x1|TOTAL|videos" gpurun_out/ab38.log | cut -c1-60
timeout 300 python bench.py --model full --steps 10 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/bench_r01_full.log 2>&1; echo "bench full rc=$?"; tail -c 2900 gpurun_out/bench_r01_full.log | head -c 250
timeout 300 python tools/bench_latency.py > gpurun_out/latency_r01.log 2>&1; echo "latency rc=$?"; tail -1 gpurun_out/latency_r01.log | cut -c1-400
