# round 2, run aa (1 GPU): compute-sanitizer over the peer-exchange kernel's single-GPU tests
mkdir -p gpurun_out
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_peer.py -x -q -m gpu 2>&1 | tail -6) > gpurun_out/r2aa_san_memcheck.log
(timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_peer.py -x -q -m gpu -k "identity or fused" 2>&1 | tail -6) > gpurun_out/r2aa_san_racecheck.log
(timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_peer.py -x -q -m gpu -k "identity or fused" 2>&1 | tail -6) > gpurun_out/r2aa_san_synccheck.log
tail -4 gpurun_out/r2aa_san_memcheck.log gpurun_out/r2aa_san_racecheck.log gpurun_out/r2aa_san_synccheck.log
timeout 300 ncu --set full --clock-control none -k regex:peer_allreduce_kernel -c 2 -f -o gpurun_out/r2aa_peer python -m pytest tests/test_gpu_peer.py -q -m gpu -k fused > gpurun_out/r2aa_ncu.log 2>&1
ncu -i gpurun_out/r2aa_peer.ncu-rep --page raw --csv > gpurun_out/r2aa_peer_raw.csv 2>/dev/null; tail -2 gpurun_out/r2aa_ncu.log
