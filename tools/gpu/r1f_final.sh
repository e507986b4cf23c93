# final tree: the whole GPU suite + the HBM-bound kernels timed alone
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/t_final.log
timeout 200 python tools/hbm_kernels_bench.py --out gpurun_out/hbm_kernels.json > gpurun_out/hbm.log 2>&1
cat gpurun_out/t_final.log; cut -c1-230 gpurun_out/hbm.log
