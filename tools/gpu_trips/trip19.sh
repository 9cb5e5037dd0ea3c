export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "test_stem" --timeout 300 > gpurun_out/t19_stem.log 2>&1; echo "stem rc=$?"; tail -25 gpurun_out/t19_stem.log
timeout 600 python -m pytest tests/test_gpu_eco.py -m gpu -q -k "test_eco_lite_n4_fast or test_eco_lite_n16 or test_pipelined" --timeout 300 > gpurun_out/t19_eco.log 2>&1; echo "eco rc=$?"; tail -15 gpurun_out/t19_eco.log
timeout 300 python tools/ab_bench.py --batch 32 rows=stem_rows:1 rows_nopool=stem_rows:2 old=stem_rows:0 > gpurun_out/ab19.log 2>&1; echo "ab rc=$?"; head -12 gpurun_out/ab19.log; tail -4 gpurun_out/ab19.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench19.log 2>&1; echo "bench rc=$?"; tail -c 2600 gpurun_out/bench19.log
