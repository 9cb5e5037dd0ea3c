export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 900 bash tools/run_gpu_suite.sh "ops eco" > gpurun_out/trip16_suite.log 2>&1
grep rc= gpurun_out/trip16_suite.log | tr '\n' ' '
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "halo" --timeout 200 > gpurun_out/trip16_halo.log 2>&1; echo "halo rc=$?"
for b in 32 64; do timeout 300 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/bench16_b$b.log 2>&1; echo "bench b$b rc=$?"; tail -c 700 gpurun_out/bench16_b$b.log; done
