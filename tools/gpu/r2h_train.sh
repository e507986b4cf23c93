# round 2, run h (1 GPU): training kernels with register->HBM stash stores -- parity tests, step timing, launch list
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_parity_sizes.py -m gpu -q -k "train or graph or overlapped or fit" 2>&1 | tail -12) > gpurun_out/r2h_tests.log
cat gpurun_out/r2h_tests.log
timeout 200 python tools/train_bench.py 30 > gpurun_out/r2h_train_bench.log 2>&1; tail -1 gpurun_out/r2h_train_bench.log
timeout 200 python tools/fwd_train_bench.py > gpurun_out/r2h_fwd_train.log 2>&1; tail -1 gpurun_out/r2h_fwd_train.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_train_launches_raw.csv python tools/train_bench.py 2 > gpurun_out/r2h_ncu_train.log 2>&1
timeout 300 python examples/train_spheres.py 3000 100 2>&1 | tail -3
