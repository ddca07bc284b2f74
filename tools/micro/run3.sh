python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/bench_modes.py --steps 3 --modes sw_align,nw_align 2>&1 | grep -E "tb_ms|\"ms\"" 
python tools/e2e_probe.py --steps 5
python tools/e2e_probe.py --steps 5 --threads 4
VERSALIGN_CUDA_MALLOC_TUNE=0 python tools/e2e_probe.py --steps 5 --threads 4 --modes legacy
VERSALIGN_CUDA_MALLOC_TUNE=0 python tools/e2e_probe.py --steps 5 --modes legacy
