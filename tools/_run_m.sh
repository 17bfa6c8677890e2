mkdir -p gpurun_out
timeout 300 python tools/gpu_kernel_check.py gn5d_silu_stats_v2 gn5d_concat_stats_v2 gn5d_c2_l0_mean3_stats_v2 2>&1 | tail -4
for o in 0 1 0 1; do (timeout 300 python bench.py --steps 10 --warmup 3 --no-library-baseline --no-inversion --no-cpu-baseline --option gn_stats_v2=$o --tag _gnv$o 2>gpurun_out/r02m_bench$o.err | tail -1) > gpurun_out/r02m_bench_gnv$o.json; python -c "
import json; d=json.load(open('gpurun_out/r02m_bench_gnv$o.json')); print('gn_stats_v2=$o', d['ms_per_step'], d['config']['eager_ms_per_step'], d['clocks']['sm_mhz'], {k:(v['ms_per_step']) for k,v in d['kernels'].items()})"; done
