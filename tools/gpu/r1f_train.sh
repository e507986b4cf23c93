# train step: parity tests + step timing (tools/train_bench.py, bench.py train leg)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_callers.py tests/test_pose_datasets.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/t_mlp.log
for n in 0 48; do echo "overlap_sms=$n: $(NERFB200_DW_OVERLAP_SMS=$n timeout 300 python tools/train_bench.py 30 2>&1 | tail -1)"; done > gpurun_out/train_overlap.log
timeout 600 python bench.py --no-cpu --no-e2e --steps 10 > gpurun_out/bench_train.json 2>gpurun_out/bench_train.err
cat gpurun_out/t_mlp.log gpurun_out/train_overlap.log; python -c "
import json; d=json.load(open('gpurun_out/bench_train.json')); print(d['value'], d['train'])"
