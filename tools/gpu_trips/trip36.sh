export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 400 python tools/ab_bench.py --batch 32 --iters 1 print=pair:0,debug_flags:16 > gpurun_out/ab36.log 2>&1; echo "ab rc=$?"; grep "persistent cta0" gpurun_out/ab36.log | tail -92 | head -92 | cut -c1-200
