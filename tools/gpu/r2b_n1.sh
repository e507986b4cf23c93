# round 2, run b (1 GPU): GPU suite with the new parity tests, the restructured bench line, ncu evidence of the shipped kernels
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2b_tests.log
cat gpurun_out/r2b_tests.log
timeout 900 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
cut -c1-1500 gpurun_out/r2b_bench_n1.json; tail -5 gpurun_out/r2b_bench_n1.err
timeout 600 python bench.py --precision tf32 --no-train --no-cpu --no-extra > gpurun_out/r2b_bench_n1_tf32.json 2> gpurun_out/r2b_bench_n1_tf32.err
cut -c1-600 gpurun_out/r2b_bench_n1_tf32.json; tail -3 gpurun_out/r2b_bench_n1_tf32.err
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2b_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-train --no-extra --no-hbm > gpurun_out/r2b_ncu_bench.log 2>&1
# ncu --set full: one fine launch of the pair kernel + its split launch, the tf32 kernel, the HBM-bound kernels
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:mlp_tc_forward_pair_kernel' -s 2 -c 2 -f -o gpurun_out/r2b_mlp_pair python tools/mlp_fine_launch.py bf16 > gpurun_out/r2b_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:mlp_tf32_forward_kernel' -s 1 -c 1 -f -o gpurun_out/r2b_mlp_tf32 python tools/mlp_fine_launch.py tf32 > gpurun_out/r2b_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:composite_fwd|sample_fine|sample_coarse' -c 8 -f -o gpurun_out/r2b_hbm_kernels python tools/hbm_kernels_bench.py --iters 1 --warmup 0 --sets 1 > gpurun_out/r2b_ncu3.log 2>&1
for r in r2b_mlp_pair r2b_mlp_tf32 r2b_hbm_kernels; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -20; tail -2 gpurun_out/r2b_ncu1.log gpurun_out/r2b_ncu2.log gpurun_out/r2b_ncu3.log
