export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_umma_persistent -s 32 -c 32 -o gpurun_out/prof_convs_r01b python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_full4.log 2>&1; echo "ncu convs rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_cl -s 4 -c 4 -o gpurun_out/prof_pools_r01b python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_pool4.log 2>&1; echo "ncu pools rc=$?"
