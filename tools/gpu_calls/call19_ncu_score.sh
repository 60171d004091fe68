# ncu --set full of the scoring head (ncu matches kernel BASE names, so the window is positioned by launch counts:
# one scoring pass = tokenize_pad + 129 gemm_tcgen05_kernel + score_finish = 131 matching launches)
K='regex:gemm_tcgen05_kernel|score_finish_kernel|tokenize_pad_kernel'
ncu --set full --clock-control none -k "$K" -s 390 -c 4 -o gpurun_out/r02_ncu_score_head python tools/ncu_targets_r02.py score > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none -k "$K" -s 295 -c 1 -o gpurun_out/r02_ncu_rope_gemm python tools/ncu_targets_r02.py score > gpurun_out/ncu_f.log 2>&1
python -m pytest tests/test_gpu_frontend.py -m gpu -q 2>&1 | tail -3
ls -la gpurun_out/r02_ncu_score_head.ncu-rep gpurun_out/r02_ncu_rope_gemm.ncu-rep
