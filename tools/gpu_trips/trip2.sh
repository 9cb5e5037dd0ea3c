bash tools/run_gpu_suite.sh "ops eco" > gpurun_out/trip2_suite.log 2>&1
export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
timeout 600 python bench.py --steps 10 --warmup 3 --batch 32 > gpurun_out/bench_b32.log 2>&1; echo "bench32 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.log 2>&1; echo "bench8 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/bench_b32_nograph.log 2>&1; echo "bench32ng rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 51 -c 3 -o gpurun_out/prof_conv_r01 python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/bench_b32.log gpurun_out/smoke.log
