# round 2, run s (8 GPUs): the data-parallel step under every schedule (NCCL / peer kernel: NVLS, unicast, IPC; eager / graph)
# and the exchange alone
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/dp_bench.py 200 > gpurun_out/r2s_dp_n8.json 2> gpurun_out/r2s_dp_n8.err
echo "dp rc=$?"; grep '^{' gpurun_out/r2s_dp_n8.json | cut -c1-200; tail -5 gpurun_out/r2s_dp_n8.err | cut -c1-300
