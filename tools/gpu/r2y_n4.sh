# round 2, run y (4 GPUs): the full bench line at N=4 on the final tree
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2y_bench_n4.json 2> gpurun_out/r2y_bench_n4.err
echo "bench rc=$?"; grep '^{' gpurun_out/r2y_bench_n4.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value']); print(json.dumps(d['train'])[:1800])"; tail -3 gpurun_out/r2y_bench_n4.err | cut -c1-300
