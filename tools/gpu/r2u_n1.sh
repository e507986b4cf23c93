# round 2, run u (1 GPU): GPU suite after the Adam helper / reduce changes, train step timings, launch list of the 512-ray step
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2u_tests.log
cat gpurun_out/r2u_tests.log
for B in 512 4096; do timeout 100 python tools/train_bench.py 50 bf16 $B 2>&1 | tail -1; timeout 100 python tools/train_bench.py 50 bf16 $B graph 2>&1 | tail -1; done > gpurun_out/r2u_steps.log
cat gpurun_out/r2u_steps.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 160 --csv --log-file gpurun_out/r2u_train512_raw.csv python tools/train_bench.py 6 bf16 512 > /dev/null 2>&1
python tools/summarize_ncu.py r2u_train512 gpurun_out/r2u_train512_raw.csv; cp profiles/r2u_train512_launches.csv gpurun_out/; head -30 gpurun_out/r2u_train512_launches.csv | cut -c1-150
