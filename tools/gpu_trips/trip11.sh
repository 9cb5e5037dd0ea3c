export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 800 bash tools/run_gpu_suite.sh "ops eco" > gpurun_out/trip11_suite.log 2>&1
grep rc= gpurun_out/trip11_suite.log | tr '\n' ' '
timeout 400 python tools/ab_bench.py --batch 32 halo= nohalo=halo:0 nodual=halo:0,dual_m:0 > gpurun_out/ab11_b32.log 2>&1; echo "ab rc=$?"; tail -48 gpurun_out/ab11_b32.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/bench11_b32.log 2>&1; tail -c 600 gpurun_out/bench11_b32.log
