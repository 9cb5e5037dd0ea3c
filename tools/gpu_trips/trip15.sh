export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
ECO_B200_LIB="$PWD/eco-efficient-video-understanding_b200/lib_old/libeco_b200_42db707.so" timeout 300 python tools/ab_bench.py --batch 32 old_staged=halo:0 old_direct=halo:0,epi_staged:0 > gpurun_out/ab15_old.log 2>&1; echo "old rc=$?"
timeout 300 python tools/ab_bench.py --batch 32 new=halo:0 new_again=halo:0 > gpurun_out/ab15_new.log 2>&1; echo "new rc=$?"
paste <(cut -c1-58 gpurun_out/ab15_old.log) <(cut -c35-58 gpurun_out/ab15_new.log) | tail -46
