export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 400 python tools/ab_bench.py --batch 32 --iters 3 base=pair:0 nostore=pair:0,debug_flags:1 noldtm=pair:0,debug_flags:3 nostore_notma=pair:0,debug_flags:9 > gpurun_out/ab35.log 2>&1; echo "ab rc=$?"; grep -E "^op|reduce|1x1|pool|conv1|TOTAL" gpurun_out/ab35.log | cut -c1-110
