# round 2, run f (2 GPUs): the failing 2-rank tests with full output; a short N=2 bench to check the exit path
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -80) > gpurun_out/r2f_multi.log
(timeout 200 python -m pytest "tests/test_gpu_mlp.py::test_graphed_train_step_equals_eager" -m gpu -q 2>&1 | tail -60) > gpurun_out/r2f_graph.log
cut -c1-220 gpurun_out/r2f_multi.log | tail -60; cut -c1-220 gpurun_out/r2f_graph.log | tail -40
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra --no-hbm > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/r2f_bench_n2.json; tail -3 gpurun_out/r2f_bench_n2.err
