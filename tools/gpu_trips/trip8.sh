export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 120 ./build/probe_umma_shift > gpurun_out/probe_shift.log 2>&1; echo "probe rc=$?"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "pooling or pool_variants" --timeout 200 > gpurun_out/trip8_pool.log 2>&1; echo "pool tests rc=$?"
timeout 600 python tools/ab_bench.py --batch 32 base= direct=epi_staged:0 nodual=dual_m:0 nodual_direct=dual_m:0,epi_staged:0 nonpersist=persistent:0 > gpurun_out/ab_b32.log 2>&1; echo "ab rc=$?"
cat gpurun_out/probe_shift.log | tail -30
