# round 2, run k (8 GPUs): schedules of the data-parallel training step (tools/dp_bench.py), default NCCL and NCCL_MAX_CTAS=4
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/dp_bench.py 100 > gpurun_out/r2k_dp_n8.json 2> gpurun_out/r2k_dp_n8.err
echo "rc=$?"; cat gpurun_out/r2k_dp_n8.json; tail -3 gpurun_out/r2k_dp_n8.err
NCCL_MAX_CTAS=4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/dp_bench.py 100 > gpurun_out/r2k_dp_n8_ctas4.json 2> gpurun_out/r2k_dp_n8_ctas4.err
echo "rc=$?"; cat gpurun_out/r2k_dp_n8_ctas4.json; tail -3 gpurun_out/r2k_dp_n8_ctas4.err
