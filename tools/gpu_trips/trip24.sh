export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "test_stem" --timeout 300 > gpurun_out/t24_stem.log 2>&1; echo "stem rc=$?"; tail -25 gpurun_out/t24_stem.log
timeout 900 python -m pytest tests/test_gpu_eco.py -m gpu -q --timeout 600 > gpurun_out/t24_eco.log 2>&1; echo "eco rc=$?"; tail -12 gpurun_out/t24_eco.log
timeout 400 python tools/ab_bench.py --batch 32 --iters 3 base= print=debug_flags:16 cells=stem_direct:0 > gpurun_out/ab24.log 2>&1; echo "ab rc=$?"; grep -m 4 "stem_rows cta0" gpurun_out/ab24.log; grep -A6 "^op " gpurun_out/ab24.log; tail -3 gpurun_out/ab24.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench24.log 2>&1; echo "bench rc=$?"; tail -c 2600 gpurun_out/bench24.log
