mkdir -p gpurun_out
timeout 400 python tools/gpu_kernel_check.py self_attn_d40_ptmem self_attn_d40_ptmem_ragged self_attn_d80_ptmem cross_attn_d40_ptmem cross_attn_d80_ptmem > gpurun_out/r02k_checks.log 2>&1; tail -7 gpurun_out/r02k_checks.log
timeout 300 python tools/gpu_attn_ab.py --reps 7 attn_persist=1 attn_persist=1:attn_ptmem=1 attn_persist=1:attn_ptmem=1:attn_poly=0 attn_persist=1:attn_poly=0 2>&1 | grep -v debug | tail -18 | tee gpurun_out/r02k_attn_ab.log
