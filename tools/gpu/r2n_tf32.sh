# round 2, run n (1 GPU): tf32 kernel with the bias on the tensor core + 4 weight stages: parity tests, bench line, ncu
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_parity_sizes.py -m gpu -q -k "tf32" 2>&1 | tail -6) > gpurun_out/r2n_tests.log
cat gpurun_out/r2n_tests.log
timeout 600 python bench.py --precision tf32 --no-train --no-cpu --no-extra --no-hbm > gpurun_out/r2n_bench_n1_tf32.json 2> gpurun_out/r2n_bench_n1_tf32.err
cut -c1-300 gpurun_out/r2n_bench_n1_tf32.json; tail -3 gpurun_out/r2n_bench_n1_tf32.err
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:mlp_tf32_forward_kernel' -s 1 -c 1 -f -o gpurun_out/r2n_mlp_tf32 python tools/mlp_fine_launch.py tf32 > gpurun_out/r2n_ncu.log 2>&1
ncu -i gpurun_out/r2n_mlp_tf32.ncu-rep --page raw --csv > gpurun_out/r2n_mlp_tf32_raw.csv 2>/dev/null; tail -1 gpurun_out/r2n_ncu.log
