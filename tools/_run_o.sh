mkdir -p gpurun_out
timeout 400 python tools/gpu_kernel_check.py gn_frame_fused_l1 gn_frame_fused_l2 gn_frame_fused_l3_silu gn_frame_fused_odd gn_frame_two_kernels gn_frame_too_large_for_cluster gn_frame_l0 gn_frame gn5d_silu 2>&1 | tail -11
for o in 0 1 0 1; do (timeout 300 python bench.py --steps 10 --warmup 3 --no-library-baseline --no-inversion --no-cpu-baseline --option gn_fused=$o --tag _gnf$o 2>gpurun_out/r02o_bench$o.err | tail -1) > gpurun_out/r02o_bench_gnf$o.json; python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_gnf$o.json')); print('gn_fused=$o', d['ms_per_step'], d['config']['eager_ms_per_step'], d['clocks']['sm_mhz'], d['finite'], {k:(v['ms_per_step'],v['launches']) for k,v in d['kernels'].items()})"; tail -2 gpurun_out/r02o_bench$o.err; done
