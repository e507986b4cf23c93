# round 2, run l (2 GPUs): data-parallel step schedules incl. the captured fork/join all-reduce; 2-rank parity test
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/dp_bench.py 60 > gpurun_out/r2l_dp_n2.json 2> gpurun_out/r2l_dp_n2.err
echo "rc=$?"; cat gpurun_out/r2l_dp_n2.json; tail -3 gpurun_out/r2l_dp_n2.err
(timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2l_multi.log; cat gpurun_out/r2l_multi.log
