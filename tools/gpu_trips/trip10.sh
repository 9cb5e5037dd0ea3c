export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "halo or test_conv2d or inception" --timeout 300 > gpurun_out/trip10_halo.log 2>&1; echo "halo tests rc=$?"; tail -5 gpurun_out/trip10_halo.log
timeout 400 python tools/ab_bench.py --batch 32 halo= nohalo=halo:0 > gpurun_out/ab10_b32.log 2>&1; echo "ab rc=$?"; tail -60 gpurun_out/ab10_b32.log
