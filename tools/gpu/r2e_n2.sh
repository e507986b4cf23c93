# round 2, run e (2 GPUs): 2-rank NCCL parity test, device-binding test, bench line at N=2 (strong-scaling view + gather, DP train)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_mlp.py::test_context_is_bound_to_its_device" "tests/test_gpu_mlp.py::test_graphed_train_step_equals_eager" -m gpu -q 2>&1 | tail -15) > gpurun_out/r2e_tests.log
cat gpurun_out/r2e_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
cut -c1-400 gpurun_out/r2e_bench_n2.json; tail -5 gpurun_out/r2e_bench_n2.err
