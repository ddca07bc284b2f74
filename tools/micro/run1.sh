set -x
lscpu | egrep "Model name|^CPU\(s\)|Thread|Core|Socket|NUMA|MHz|L3" 
free -g | head -2
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv
./tools/micro/pipe_mix
for T in 4 16; do ./tools/micro/malloc_probe $T | tail -3; ./tools/micro/malloc_probe $T tuned | tail -3; done
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python tools/bench_modes.py --steps 3 2>&1 | tail -40
