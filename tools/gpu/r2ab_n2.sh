# round 2, run ab (2 GPUs): experiment -- the exchange alone with the per-CTA fence at gpu scope (variant 1) vs sys scope (0); the switch was removed after the measurement, gpu scope is what the kernel does
mkdir -p gpurun_out
for v in 0 1; do
NERFB200_PEER_VARIANT=$v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$v tools/dp_bench.py 10 exchange > gpurun_out/r2ab_exch_n2_v$v.json 2> gpurun_out/r2ab_exch_n2_v$v.err
echo "variant $v rc=$?"; grep '^{' gpurun_out/r2ab_exch_n2_v$v.json | cut -c1-2500
done
