export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
bash tools/run_gpu_suite.sh "ops eco" 2>&1 | grep -E "rc=|passed|failed|error" | head -60
for i in 1 2; do
ECO_B200_LIB=$PWD/eco-efficient-video-understanding_b200/lib_old/libeco_b200_pre256.so timeout 200 python tools/ab_bench.py --batch 32 base= > gpurun_out/ab18_old_$i.log 2>&1; echo "ab old rc=$?"; tail -3 gpurun_out/ab18_old_$i.log
timeout 200 python tools/ab_bench.py --batch 32 base= > gpurun_out/ab18_new_$i.log 2>&1; echo "ab new rc=$?"; tail -3 gpurun_out/ab18_new_$i.log
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench18.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench18.log
