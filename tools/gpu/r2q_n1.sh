# round 2, run q (1 GPU): GPU suite + smoke + bench line with the one-call forward
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2q_tests.log
cat gpurun_out/r2q_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --no-extra > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err
cut -c1-300 gpurun_out/r2q_bench_n1.json; tail -3 gpurun_out/r2q_bench_n1.err
