export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "multicast" --timeout 120 > gpurun_out/t32_mc.log 2>&1; echo "multicast tests rc=$?"; tail -25 gpurun_out/t32_mc.log | cut -c1-300
timeout 300 python tools/ab_bench.py --batch 32 --iters 3 base=multicast:0 mc1=multicast:1 mc2=multicast:2 > gpurun_out/ab32.log 2>&1; echo "ab rc=$?"; grep -A40 "^op " gpurun_out/ab32.log | cut -c1-100
