python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python tools/bench_modes.py --steps 3 --modes sw_align,nw_align 2>&1 | grep -E "ms|gcups" 
