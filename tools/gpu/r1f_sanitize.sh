# compute-sanitizer over the HBM-bound kernels' parity tests (new shared-memory code paths of this round)
mkdir -p gpurun_out
SEL='fine_sampler or integrator or coarse_sampler'
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "$SEL" 2>&1 | tail -12) > gpurun_out/san_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "$SEL" 2>&1 | tail -12) > gpurun_out/san_racecheck.log
(timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "$SEL" 2>&1 | tail -12) > gpurun_out/san_synccheck.log
(timeout 300 python examples/train_spheres.py 3000 100 2>&1 | tail -4) > gpurun_out/train_spheres.log
tail -4 gpurun_out/san_memcheck.log gpurun_out/san_racecheck.log gpurun_out/san_synccheck.log gpurun_out/train_spheres.log
