export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "conv or residual or inception or stem" --timeout 300 > gpurun_out/trip13_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/trip13_ops.log
timeout 400 python tools/ab_bench.py --batch 32 base= nohalo=halo:0 nostore=halo:0,debug_flags:1 noldtm=halo:0,debug_flags:3 nomma=halo:0,debug_flags:7 > gpurun_out/ab13_b32.log 2>&1; echo "ab rc=$?"; tail -48 gpurun_out/ab13_b32.log
