mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_stencil2_gpu.py tests/test_stencil_pair_gpu.py -m gpu -x -q > gpurun_out/tile_pytest.txt 2>&1; tail -3 gpurun_out/tile_pytest.txt
timeout 300 python tools/bench_metric_stencils.py > gpurun_out/r2_metric_stencils_final.txt 2>&1; cat gpurun_out/r2_metric_stencils_final.txt
timeout 300 python tools/bench_pair.py 2>&1 | head -2
