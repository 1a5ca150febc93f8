mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_stencil2_gpu.py -m gpu -x -q -k "multi" > gpurun_out/multi_pytest.txt 2>&1; tail -3 gpurun_out/multi_pytest.txt | cut -c1-220
rm -f gpurun_out/r2_multi_tma_sweep.txt
for v in "XG_MULTI_TMA=1" "XG_MULTI_CTAS=4" "XG_MULTI_CTAS=4 XG_MULTI_NST=1" "XG_MULTI_NST=3" "XG_MULTI_TMA=0"; do
  env $v timeout 200 python tools/bench_multi.py "$v" >> gpurun_out/r2_multi_tma_sweep.txt 2>&1
done
cat gpurun_out/r2_multi_tma_sweep.txt
