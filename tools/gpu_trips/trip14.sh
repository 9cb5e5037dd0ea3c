export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "halo or stem or test_conv2d or inception" --timeout 300 > gpurun_out/trip14_ops.log 2>&1; echo "ops rc=$?"; tail -15 gpurun_out/trip14_ops.log | cut -c1-300
timeout 400 python tools/ab_bench.py --batch 32 base= nohalo=halo:0 stream=halo:3 > gpurun_out/ab14_b32.log 2>&1; echo "ab rc=$?"; tail -48 gpurun_out/ab14_b32.log
