export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t34_all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 gpurun_out/t34_all_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r01.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r01_default.log 2>&1; echo "bench default rc=$?"; tail -c 2900 gpurun_out/bench_r01_default.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches34.log 2>&1; echo "ncu launches rc=$?"
