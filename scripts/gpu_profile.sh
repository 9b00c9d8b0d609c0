# full evidence run: parity tests, the contract bench line, ncu launch list, ncu --set full of the cull + pyramid kernels.  TAG names the outputs.
mkdir -p gpurun_out
TAG=${1:-rXX}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader; nproc
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cat gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json
timeout 600 python scripts/cull_microbench.py --iters 30 > gpurun_out/${TAG}_microbench.jsonl 2>&1; cat gpurun_out/${TAG}_microbench.jsonl
timeout 600 python scripts/inst_cluster_microbench.py > gpurun_out/${TAG}_inst_cluster.jsonl 2>&1; cut -c1-160 gpurun_out/${TAG}_inst_cluster.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
if [ "${2:-}" = full ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_cull_kernel -s 4 -c 1 -o gpurun_out/${TAG}_prof_late python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof_late.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:early_stream_kernel -s 4 -c 1 -o gpurun_out/${TAG}_prof_early python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof_early.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyramid_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_pyr python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_prof_pyr.log 2>&1
fi
ls -la gpurun_out | tail -12
