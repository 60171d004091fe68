# ncu --set full for the kernels outside the bench step: scoring head and generation loop
ncu --set full --clock-control none -k 'regex:gemm_tcgen05_kernel<2, 6|gemm_tcgen05_kernel<2, 5|score_finish|tokenize_pad' -s 10 -c 5 -o gpurun_out/r02_ncu_score python tools/ncu_targets_r02.py score > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none -k 'regex:sample_step|advance2|gemm_smallm|decode_attn_tma|hyena_step|rmsnorm_row' -s 400 -c 12 -o gpurun_out/r02_ncu_gen python tools/ncu_targets_r02.py gen > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/r02_ncu_score.ncu-rep gpurun_out/r02_ncu_gen.ncu-rep; tail -2 gpurun_out/ncu_d.log; tail -2 gpurun_out/ncu_e.log
