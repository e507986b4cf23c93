# GPU-box run behind profiles/r1f_*: parity tests, stand-alone HBM kernel timings, the bench line, ncu launch lists.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/t1.log
timeout 300 python tools/hbm_kernels_bench.py --out gpurun_out/hbm_kernels.json > gpurun_out/hbm.log 2>&1
timeout 900 python bench.py > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1f_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-train > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1f_train_launches_raw.csv python tools/train_bench.py 2 > gpurun_out/ncu_train.log 2>&1
tail -4 gpurun_out/t1.log; cat gpurun_out/hbm.log; cat gpurun_out/r1f_bench_n1.json; tail -3 gpurun_out/r1f_bench_n1.err; tail -2 gpurun_out/ncu_train.log
