mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest_gpu.log
cat gpurun_out/r1_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
cat gpurun_out/r1_bench_n1.json; tail -5 gpurun_out/r1_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_launches_bench.log 2>&1
tail -30 gpurun_out/r1_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:draw_cull_kernel -s 7 -c 2 -o gpurun_out/r1_prof_cull python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_prof_cull.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyramid_kernel -s 3 -c 1 -o gpurun_out/r1_prof_pyr python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_prof_pyr.log 2>&1
ls -la gpurun_out
