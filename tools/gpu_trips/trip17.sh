export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r01_default.log 2>&1; echo "bench default rc=$?"; tail -c 2500 gpurun_out/bench_r01_default.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.log 2>&1; echo "bench ref rc=$?"; tail -c 600 gpurun_out/bench_r01_reference.log
timeout 300 python bench.py --model full --steps 10 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/bench_r01_full.log 2>&1; echo "bench full rc=$?"; tail -c 900 gpurun_out/bench_r01_full.log
timeout 300 python tools/bench_latency.py > gpurun_out/latency_r01.log 2>&1; echo "latency rc=$?"; tail -2 gpurun_out/latency_r01.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches17.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_umma" -s 32 -c 32 -o gpurun_out/prof_convs_r01_final python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_full17.log 2>&1; echo "ncu full rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r01.log
