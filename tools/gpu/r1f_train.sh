# train step: parity tests + step timing (tools/train_bench.py, tools/fwd_train_bench.py, bench.py train leg)
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_mlp.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/t_mlp.log
echo "train_bench: $(timeout 120 python tools/train_bench.py 30 2>&1 | tail -1)" > gpurun_out/train_overlap.log
timeout 120 python tools/fwd_train_bench.py >> gpurun_out/train_overlap.log 2>&1
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 > gpurun_out/bench_train.json 2>gpurun_out/bench_train.err
cat gpurun_out/t_mlp.log gpurun_out/train_overlap.log; python -c "
import json; d=json.load(open('gpurun_out/bench_train.json')); print(d['value'], d['train'])"
