export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 700 bash tools/run_gpu_suite.sh "ops eco" > gpurun_out/trip6_suite.log 2>&1
grep rc= gpurun_out/trip6_suite.log
for b in 32; do timeout 300 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/bench6_b$b.log 2>&1; echo "bench b$b rc=$?"; tail -c 1000 gpurun_out/bench6_b$b.log; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches6.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pool|stem_s2d" -s 5 -c 5 -o gpurun_out/prof_pools_r01d python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_pool6.log 2>&1; echo "ncu pools rc=$?"
