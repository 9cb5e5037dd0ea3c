export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 300 python tools/ab_bench.py --batch 32 --iters 3 base=pair:0 notma=pair:0,debug_flags:8 notma_nost=pair:0,debug_flags:9 nomma=pair:0,debug_flags:4 > gpurun_out/ab31.log 2>&1; echo "ab rc=$?"; grep -A40 "^op " gpurun_out/ab31.log | cut -c1-100
