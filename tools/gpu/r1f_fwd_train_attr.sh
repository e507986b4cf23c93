# attribution of the training forward's slowdown (debug instantiation; see tools/fwd_train_bench.py)
mkdir -p gpurun_out
for m in 1 3 4 5; do timeout 120 python tools/fwd_train_bench.py $m 2>&1 | grep -E "debug mode|cta0 (mma|epi0|producer)" | tail -7; done > gpurun_out/fwd_train_attr.log
timeout 120 python tools/fwd_train_bench.py >> gpurun_out/fwd_train_attr.log 2>&1
cat gpurun_out/fwd_train_attr.log
