# round 2, run d (1 GPU): accuracy of the split launch at the GPU's own sample positions, the GPU suite, smoke
mkdir -p gpurun_out
timeout 600 python tools/split_diag.py > gpurun_out/r2d_split_diag.log 2>&1; cat gpurun_out/r2d_split_diag.log | cut -c1-400
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/r2d_tests.log
cat gpurun_out/r2d_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
