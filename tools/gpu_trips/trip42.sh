export PYTHONPATH="$PWD:$PWD/tools:$PWD/eco-efficient-video-understanding_b200:$PWD/tests"
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches42.log 2>&1; echo "ncu launches rc=$?"
timeout 260 ncu --set full --clock-control none --import-source on -k regex:"conv_umma|stem_rows|pool2d_rows" -s 29 -c 29 -f -o gpurun_out/prof_convs_r01_final python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-graph > gpurun_out/ncu_full42.log 2>&1; echo "ncu full rc=$?"
