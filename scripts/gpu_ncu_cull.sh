# ncu --set full of the early + late cull kernels in steady state (bench workload), plus the launch list
mkdir -p gpurun_out
TAG=${1:-x}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:draw_cull_kernel -s 7 -c 2 -o gpurun_out/${TAG}_prof_cull python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof_cull.log 2>&1
tail -3 gpurun_out/${TAG}_prof_cull.log
