mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r02g_pytest.log; tail -4 gpurun_out/r02g_pytest.log
timeout 300 python tools/gpu_attn_ab.py --reps 7 attn_persist=1 attn_persist=1:attn_pingpong=0 attn_persist=1:attn_pingpong=0:attn_poly=2 attn_persist=1:attn_poly=1 2>&1 | tail -20 | tee gpurun_out/r02g_attn_ab.log
for o in 1 0; do (timeout 300 python bench.py --steps 8 --warmup 3 --no-library-baseline --no-inversion --no-cpu-baseline --option subpixel=$o --tag _subpixel$o 2>gpurun_out/r02g_bench$o.err | tail -1) > gpurun_out/r02g_bench_subpixel$o.json; python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_subpixel$o.json')); print('subpixel=$o', d['ms_per_step'], d['config']['eager_ms_per_step'], {k:(v['ms_per_step'],v['launches']) for k,v in d['kernels'].items()})"; done
