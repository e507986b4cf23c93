# round 2, run w (2 GPUs): the 2-rank tests on the final tree and the full bench line at N=2 (peer-memory gradient exchange)
mkdir -p gpurun_out
(timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -12) > gpurun_out/r2w_multi.log
cut -c1-250 gpurun_out/r2w_multi.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err
echo "bench rc=$?"; grep '^{' gpurun_out/r2w_bench_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value']); print(json.dumps(d['train'])[:1500])"; tail -3 gpurun_out/r2w_bench_n2.err | cut -c1-300
