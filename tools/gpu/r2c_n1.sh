# round 2, run c (1 GPU): the whole GPU suite (new parity / graph / split tests), bench lines, train-step launch list
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/r2c_tests.log
cat gpurun_out/r2c_tests.log
timeout 900 python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
cut -c1-300 gpurun_out/r2c_bench_n1.json; tail -5 gpurun_out/r2c_bench_n1.err
timeout 600 python bench.py --precision tf32 --no-train --no-cpu --no-extra --no-hbm > gpurun_out/r2c_bench_n1_tf32.json 2> gpurun_out/r2c_bench_n1_tf32.err
cut -c1-300 gpurun_out/r2c_bench_n1_tf32.json; tail -3 gpurun_out/r2c_bench_n1_tf32.err
timeout 300 python tools/train_bench.py 20 > gpurun_out/r2c_train_bench.log 2>&1; tail -2 gpurun_out/r2c_train_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2c_train_launches_raw.csv python tools/train_bench.py 2 > gpurun_out/r2c_ncu_train.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
