# round 2, run i (1 GPU): one rank's share of the 8-GPU data-parallel step (512 rays): step time eager/graph + launch list
mkdir -p gpurun_out
for B in 512 1024 4096; do timeout 100 python tools/train_bench.py 50 bf16 $B 2>&1 | tail -1; timeout 100 python tools/train_bench.py 50 bf16 $B graph 2>&1 | tail -1; done > gpurun_out/r2i_steps.log
cat gpurun_out/r2i_steps.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2i_train512_launches_raw.csv python tools/train_bench.py 2 bf16 512 > gpurun_out/r2i_ncu.log 2>&1
tail -2 gpurun_out/r2i_ncu.log
