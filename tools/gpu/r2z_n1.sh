# round 2, run z (1 GPU): the GPU suite on the final tree, smoke, the default bench line, train step timings
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/r2z_tests.log
cut -c1-250 gpurun_out/r2z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for B in 512 4096; do timeout 100 python tools/train_bench.py 50 bf16 $B 2>&1 | tail -1; timeout 100 python tools/train_bench.py 50 bf16 $B graph 2>&1 | tail -1; done > gpurun_out/r2z_steps.log
cat gpurun_out/r2z_steps.log
timeout 900 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2z_bench_n1.json; tail -3 gpurun_out/r2z_bench_n1.err
