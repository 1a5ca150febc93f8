mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_stencil2_gpu.py tests/test_stencil_pair_gpu.py -m gpu -x -q > gpurun_out/tile_pytest.txt 2>&1; tail -3 gpurun_out/tile_pytest.txt
rm -f gpurun_out/r2_tile_sweep.txt
for v in "XG_TILE_TMA=1" "XG_TILE_CTAS=3 XG_TILE_NST=1" "XG_TILE_CTAS=3 XG_TILE_NST=2" "XG_TILE_CTAS=3 XG_TILE_NST=3" "XG_TILE_CTAS=2 XG_TILE_NST=2" "XG_TILE_CTAS=2 XG_TILE_NST=3" "XG_TILE_CTAS=2 XG_TILE_NST=4" "XG_TILE_TMA=0"; do
  env $v timeout 200 python tools/bench_tile.py "$v" >> gpurun_out/r2_tile_sweep.txt 2>&1
done
cat gpurun_out/r2_tile_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_stencil -c 1 -o gpurun_out/r2_derivative_y_tile python tools/prof_one.py derivative_y > gpurun_out/tile_ncu.log 2>&1; tail -2 gpurun_out/tile_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_stencil -c 1 -o gpurun_out/r2_divergence_tile python tools/prof_one.py divergence > gpurun_out/tile_ncu2.log 2>&1; tail -2 gpurun_out/tile_ncu2.log
