# ncu --set full of one kernel (regex $1) in the bench workload; outputs gpurun_out/$2_prof.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-4} -c 1 -o gpurun_out/$2_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/$2_prof.log 2>&1
tail -2 gpurun_out/$2_prof.log | cut -c1-300
