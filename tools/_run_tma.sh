mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stencil2_gpu.py -m gpu -x -q -k "shared_divisor or metric" > gpurun_out/tma_pytest.txt 2>&1; tail -3 gpurun_out/tma_pytest.txt
rm -f gpurun_out/r2_row_tma_sweep.txt
for c in 2 3 4; do for n in 1 2 3; do for h in 1 0; do
  v="XG_ROW_TMA_CTAS=$c XG_ROW_TMA_NST=$n XG_ROW_TMA_HINT=$h"
  env $v timeout 200 python tools/bench_row_tma.py "$v" >> gpurun_out/r2_row_tma_sweep.txt 2>&1
done; done; done
for v in "XG_ROW_TMA_U=8 XG_ROW_TMA_CTAS=2 XG_ROW_TMA_NST=2" "XG_ROW_TMA_U=8 XG_ROW_TMA_CTAS=3 XG_ROW_TMA_NST=1" "XG_ROW_TMA_U=8 XG_ROW_TMA_CTAS=3 XG_ROW_TMA_NST=2" "XG_ROW_TMA_RB=64" "XG_ROW_TMA_RB=256" "XG_ROW_TMA_RB=512" "XG_ROW_TMA=0"; do
  env $v timeout 200 python tools/bench_row_tma.py "$v" >> gpurun_out/r2_row_tma_sweep.txt 2>&1
done
cat gpurun_out/r2_row_tma_sweep.txt
