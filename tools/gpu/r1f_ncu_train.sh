mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:mlp_tc_forward_pair_kernel|bwd_data_pair_kernel' -s 8 -c 4 -f -o gpurun_out/train_r1f python tools/train_bench.py 3 > gpurun_out/ncu_train_full.log 2>&1
tail -3 gpurun_out/ncu_train_full.log
