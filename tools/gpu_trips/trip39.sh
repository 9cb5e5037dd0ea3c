export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_r01_n4.log 2>&1; echo "bench n4 rc=$?"; grep "^{" gpurun_out/bench_r01_n4.log | cut -c1-700
