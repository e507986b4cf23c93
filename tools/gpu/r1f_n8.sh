# 8-GPU bench line (weak scaling: one view per GPU; data-parallel train step at global batch 4096)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r1f_bench_n8.json 2> gpurun_out/r1f_bench_n8.err
cut -c1-300 gpurun_out/r1f_bench_n8.json; tail -3 gpurun_out/r1f_bench_n8.err
