export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 800 bash tools/run_gpu_suite.sh "ops eco" > gpurun_out/trip9_suite.log 2>&1
grep rc= gpurun_out/trip9_suite.log | tr '\n' ' '
timeout 400 python tools/ab_bench.py --batch 32 auto= always=persistent:2 never=persistent:0 > gpurun_out/ab9_b32.log 2>&1; echo "ab rc=$?"; tail -5 gpurun_out/ab9_b32.log
for b in 32 64; do timeout 300 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/bench9_b$b.log 2>&1; echo "bench b$b rc=$?"; tail -c 1100 gpurun_out/bench9_b$b.log; done
