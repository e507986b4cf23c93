# round 2, run af (1 GPU): the GPU suite with the fused training integrator (nerfb200_composite_train); train step timings
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2af_tests.log
cut -c1-220 gpurun_out/r2af_tests.log | tail -40
(timeout 60 python tools/train_bench.py 50 bf16 512 graph 2>&1 | tail -1; timeout 60 python tools/train_bench.py 50 bf16 4096 graph 2>&1 | tail -1) > gpurun_out/r2af_steps.log
cat gpurun_out/r2af_steps.log
