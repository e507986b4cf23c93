# GPU-box check of the HBM-bound kernels: parity tests, stand-alone timings, one ncu --set full capture.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/t1.log
timeout 300 python tools/hbm_kernels_bench.py --out gpurun_out/hbm_kernels.json > gpurun_out/hbm.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:composite_fwd|sample_fine' -c 8 -f -o gpurun_out/hbm_r1f python tools/hbm_kernels_bench.py --iters 1 --warmup 0 --sets 1 > gpurun_out/ncu.log 2>&1
tail -5 gpurun_out/t1.log; cat gpurun_out/hbm.log; tail -3 gpurun_out/ncu.log
