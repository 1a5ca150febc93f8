#!/bin/bash
# Short GPU-box session: full parity suite, smoke, one driver-style bench line (tools/gpu_round.sh adds the reference arm and ncu).
TAG=r2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.txt
python bench.py --steps 20 --warmup 5 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['ms_per_step'])
for e in d['extra']: print(e['config'][:80], e.get('ms', e.get('ms_per_step_ops_only')), e.get('frac_of_peak', e.get('frac_of_peak_ops_only')), e.get('error'))"
