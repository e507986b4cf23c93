# final tree: the whole GPU suite + the HBM-bound kernels at the cfg5 shape (128 + 256 samples)
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/t_final.log
timeout 200 python tools/hbm_kernels_bench.py --nc 128 --nf 256 --out gpurun_out/hbm_kernels_128_256.json > gpurun_out/hbm_128_256.log 2>&1
cat gpurun_out/t_final.log; cut -c1-260 gpurun_out/hbm_128_256.log
