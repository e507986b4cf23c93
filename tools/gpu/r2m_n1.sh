# round 2, run m (1 GPU): final-tree GPU suite + smoke + default bench line
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2m_tests.log
cat gpurun_out/r2m_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
cut -c1-300 gpurun_out/r2m_bench_n1.json; tail -3 gpurun_out/r2m_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err
cut -c1-400 gpurun_out/r2m_bench_ref.json
