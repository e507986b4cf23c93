# 2-GPU check of bench.py under torchrun (render weak scaling + data-parallel train step)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1f_bench_n2.json 2> gpurun_out/r1f_bench_n2.err
cat gpurun_out/r1f_bench_n2.json | cut -c1-600; tail -5 gpurun_out/r1f_bench_n2.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
