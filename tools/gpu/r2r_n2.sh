# round 2, run r (2 GPUs): the peer-memory gradient exchange -- 2-rank parity tests, then the data-parallel step under
# every schedule (NCCL / peer kernel, eager / graph)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2r_topo.txt 2>&1
(timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -60) > gpurun_out/r2r_multi.log
cut -c1-250 gpurun_out/r2r_multi.log | tail -40
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dp_bench.py 100 > gpurun_out/r2r_dp_n2.json 2> gpurun_out/r2r_dp_n2.err
echo "dp rc=$?"; cut -c1-1500 gpurun_out/r2r_dp_n2.json; tail -5 gpurun_out/r2r_dp_n2.err | cut -c1-300
