# round 2, run p (4 GPUs): DP step schedules, with the parameter difference after ONE step between schedules
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 tools/dp_bench.py 60 > gpurun_out/r2p_dp_n4.json 2> gpurun_out/r2p_dp_n4.err
echo "rc=$?"; cat gpurun_out/r2p_dp_n4.json; tail -2 gpurun_out/r2p_dp_n4.err
