export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t25_all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 gpurun_out/t25_all_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r01.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r01_default.log 2>&1; echo "bench default rc=$?"; tail -c 2700 gpurun_out/bench_r01_default.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.log 2>&1; echo "bench ref rc=$?"; tail -c 700 gpurun_out/bench_r01_reference.log
timeout 300 python bench.py --model full --steps 10 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/bench_r01_full.log 2>&1; echo "bench full rc=$?"; tail -c 1200 gpurun_out/bench_r01_full.log
timeout 300 python tools/bench_latency.py > gpurun_out/latency_r01.log 2>&1; echo "latency rc=$?"; tail -2 gpurun_out/latency_r01.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches25.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_umma|stem_rows|pool2d_rows" -s 28 -c 28 -f -o gpurun_out/prof_convs_r01_final python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_full25.log 2>&1; echo "ncu full rc=$?"
