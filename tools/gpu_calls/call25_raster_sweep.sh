# big-tile GEMM: DRAM bytes + duration per rasterisation (grouping direction x group size) for the four 8k-step shapes
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv \
  --log-file gpurun_out/r02_raster_sweep.csv python tools/gemm_raster_sweep.py --ncu > gpurun_out/r02_raster_ncu.log 2>&1
python tools/gemm_raster_sweep.py --table gpurun_out/r02_raster_sweep.csv | tee gpurun_out/r02_raster_sweep.txt
echo "--- isolated, CUDA events, L2 flushed" | tee -a gpurun_out/r02_raster_sweep.txt
timeout 300 python tools/gemm_raster_sweep.py 2>&1 | tee -a gpurun_out/r02_raster_sweep.txt
