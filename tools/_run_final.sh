TAG=r2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.txt
python bench.py --steps 20 --warmup 5 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['ms_per_step']); print(json.dumps(d.get('extra',{}))[:3000])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 python tools/kernel_bench_all.py > gpurun_out/${TAG}_kernel_bench_all.txt 2>&1; tail -40 gpurun_out/${TAG}_kernel_bench_all.txt
