# phase-split backward: parity tests + train step timing for a few SM splits
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -8) > gpurun_out/t_mlp.log
for n in 0 32 48 64; do echo "overlap_sms=$n: $(NERFB200_DW_OVERLAP_SMS=$n timeout 300 python tools/train_bench.py 30 2>&1 | tail -1)"; done > gpurun_out/train_overlap.log
cat gpurun_out/t_mlp.log gpurun_out/train_overlap.log
