#!/bin/bash
# One GPU-box session: parity tests, smoke, both bench arms, ncu launch list + full captures of the two headline kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.txt
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/${TAG}_bench_reference.json
python bench.py 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stencil -s 18 -c 6 -f \
    -o gpurun_out/${TAG}_prof_stencil python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
