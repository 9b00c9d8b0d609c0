# trimmed evidence run (one GPU, a few minutes): parity tests, smoke, the contract bench line, ncu --set full of the late kernel, ncu launch list.  TAG names the outputs.
mkdir -p gpurun_out
TAG=${1:-rXX}
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 100 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 240 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cat gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
timeout 120 ncu --set full --clock-control none --import-source on -k regex:stream_cull_kernel -s 4 -c 1 -o gpurun_out/${TAG}_prof_late python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof_late.log 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la gpurun_out | tail -6
