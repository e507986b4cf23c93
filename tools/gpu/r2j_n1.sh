# round 2, run j (1 GPU): the GPU suite on the final tree, sanitizers over this round's new kernels, bench line, train step
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2j_tests.log
cat gpurun_out/r2j_tests.log
SEL='split or tf32 or tensor_core_path or ragged'
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mlp.py -x -q -k "$SEL" 2>&1 | tail -6) > gpurun_out/r2j_san_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_mlp.py -x -q -k "split and 300" 2>&1 | tail -6) > gpurun_out/r2j_san_racecheck.log
(timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_mlp.py -x -q -k "(split and 300) or (tf32 and 129)" 2>&1 | tail -6) > gpurun_out/r2j_san_synccheck.log
tail -3 gpurun_out/r2j_san_memcheck.log gpurun_out/r2j_san_racecheck.log gpurun_out/r2j_san_synccheck.log
for B in 512 4096; do timeout 100 python tools/train_bench.py 50 bf16 $B 2>&1 | tail -1; timeout 100 python tools/train_bench.py 50 bf16 $B graph 2>&1 | tail -1; done > gpurun_out/r2j_steps.log
cat gpurun_out/r2j_steps.log
timeout 900 python bench.py > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
cut -c1-300 gpurun_out/r2j_bench_n1.json; tail -3 gpurun_out/r2j_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:mlp_tc_forward_pair_kernel|bwd_data_pair_kernel|dw_kernel' -s 6 -c 6 -f -o gpurun_out/r2j_train python tools/train_bench.py 3 > gpurun_out/r2j_ncu_train.log 2>&1
ncu -i gpurun_out/r2j_train.ncu-rep --page raw --csv > gpurun_out/r2j_train_raw.csv 2>/dev/null; tail -2 gpurun_out/r2j_ncu_train.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
