# round 2, first GPU run: new split / tf32 kernels -- error statistics at cfg1 size, timings, then the GPU suite
mkdir -p gpurun_out
timeout 900 python tools/parity_diag.py 100 gpurun_out/r2a_parity_diag.json > gpurun_out/r2a_diag.log 2>&1
tail -40 gpurun_out/r2a_diag.log
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2a_tests.log
cat gpurun_out/r2a_tests.log
