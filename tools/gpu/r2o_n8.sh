# round 2, run o (8 GPUs): final-tree bench lines at N=8 and N=4, DP step schedules
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2o_bench_n8.json 2> gpurun_out/r2o_bench_n8.err
echo "rc=$?"; cut -c1-200 gpurun_out/r2o_bench_n8.json; tail -2 gpurun_out/r2o_bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/dp_bench.py 100 > gpurun_out/r2o_dp_n8.json 2> gpurun_out/r2o_dp_n8.err
echo "rc=$?"; cat gpurun_out/r2o_dp_n8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 --no-hbm > gpurun_out/r2o_bench_n4.json 2> gpurun_out/r2o_bench_n4.err
echo "rc=$?"; cut -c1-200 gpurun_out/r2o_bench_n4.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 --no-hbm > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err
echo "rc=$?"; cut -c1-200 gpurun_out/r2o_bench_n2.json
