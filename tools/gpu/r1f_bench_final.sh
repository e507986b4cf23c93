# final bench line + render launch list of the final tree
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1f_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-train > gpurun_out/ncu_bench.log 2>&1
cut -c1-200 gpurun_out/r1f_bench_n1.json; tail -2 gpurun_out/r1f_bench_n1.err
