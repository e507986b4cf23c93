# round 2, run x (8 GPUs): the full bench line at N=8 on the final tree (peer-memory gradient exchange in the DP train step)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2x_bench_n8.json 2> gpurun_out/r2x_bench_n8.err
echo "bench rc=$?"; grep '^{' gpurun_out/r2x_bench_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value']); print(json.dumps(d['train'])[:1800]); print(json.dumps(d['workloads'])[:600])"; tail -3 gpurun_out/r2x_bench_n8.err | cut -c1-300
