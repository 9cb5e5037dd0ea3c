export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_umma_persistent|conv_halo" -s 32 -c 6 -o gpurun_out/prof_small_r01f python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_full12.log 2>&1; echo "ncu rc=$?"
