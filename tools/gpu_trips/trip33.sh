export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "multicast or halo" --timeout 120 > gpurun_out/t33.log 2>&1; echo "multicast+halo tests rc=$?"; tail -5 gpurun_out/t33.log | cut -c1-300
timeout 300 python tools/ab_bench.py --batch 32 --iters 3 base= halo1=halo:1 halo3=halo:3 > gpurun_out/ab33.log 2>&1; echo "ab rc=$?"; grep -A40 "^op " gpurun_out/ab33.log | cut -c1-100
