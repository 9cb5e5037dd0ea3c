export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "cta_pair" --timeout 120 > gpurun_out/t29_pair.log 2>&1; echo "pair tests rc=$?"; tail -25 gpurun_out/t29_pair.log | cut -c1-300
timeout 300 python tools/ab_bench.py --batch 32 --iters 3 base= pair1=pair:1 pair2=pair:2 > gpurun_out/ab29.log 2>&1; echo "ab rc=$?"; grep -A40 "^op " gpurun_out/ab29.log | cut -c1-100; tail -5 gpurun_out/ab29.log | cut -c1-300
