# round 2, run t (8 GPUs): the gradient exchange alone, peer kernel vs NCCL
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/dp_bench.py 10 exchange > gpurun_out/r2t_exch_n8.json 2> gpurun_out/r2t_exch_n8.err
echo "rc=$?"; grep '^{' gpurun_out/r2t_exch_n8.json | cut -c1-1500; tail -5 gpurun_out/r2t_exch_n8.err | cut -c1-300
