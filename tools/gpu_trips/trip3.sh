export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
bash tools/run_gpu_suite.sh "ops eco" > gpurun_out/trip3_suite.log 2>&1
grep rc= gpurun_out/trip3_suite.log
for b in 8 32 64; do timeout 600 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/bench3_b$b.log 2>&1; echo "bench b$b rc=$?"; tail -c 1500 gpurun_out/bench3_b$b.log; done
timeout 600 python bench.py --steps 10 --warmup 3 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/bench3_b32_nograph.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches3.log 2>&1; echo "ncu launches rc=$?"
