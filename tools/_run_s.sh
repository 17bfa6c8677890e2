mkdir -p gpurun_out
timeout 300 python tools/gpu_kernel_check.py temporal_attn_d40 temporal_attn_d40_scalar_stores temporal_attn_d80_f16 temporal_attn_d160_f3 temporal_attn_f24 2>&1 | tail -6
for o in 0 1 0 1; do (timeout 300 python bench.py --steps 10 --warmup 3 --no-library-baseline --no-inversion --no-cpu-baseline --option tattn_vst=$o --tag _tv$o 2>gpurun_out/r02s_bench$o.err | tail -1) > gpurun_out/r02s_bench_tv$o.json; python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_tv$o.json')); print('tattn_vst=$o', d['ms_per_step'], d['clocks']['sm_mhz'], d['kernels']['temporal_attention'])"; done
(timeout 500 python tools/gpu_parity_c2.py 2>&1 | tail -2) | cut -c1-900
(timeout 400 python tools/gpu_p2p_flow.py 2>&1 | tail -1) | cut -c1-1500
