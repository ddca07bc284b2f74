VERSALIGN_CUDA_TRACE=1 python tools/e2e_probe.py --steps 1 --modes packed_pinned 2>&1 | tail -45
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err
tail -5 gpurun_out/bench_r2_a.err
