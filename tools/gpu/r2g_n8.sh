# round 2, run g (8 GPUs): the bench line at N=8 -- one view ray-sharded over 8 ranks + image gather, cfg4/cfg5, DP train step
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --no-hbm > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err
echo "rc=$?"; cut -c1-300 gpurun_out/r2g_bench_n8.json; tail -3 gpurun_out/r2g_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 --no-hbm --no-extra > gpurun_out/r2g_bench_n4.json 2> gpurun_out/r2g_bench_n4.err
echo "rc=$?"; cut -c1-300 gpurun_out/r2g_bench_n4.json; tail -3 gpurun_out/r2g_bench_n4.err
