mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r02p_pytest.log; tail -3 gpurun_out/r02p_pytest.log
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -6) | tee gpurun_out/r02p_smoke.log
(timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02p_bench.err | tail -1) > gpurun_out/r02p_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r02p_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['library_baseline']['ms_per_step'], d['inversion_step']['ms_per_step'], {k:(v['ms_per_step'],v['launches']) for k,v in d['kernels'].items()})"
(timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1) | cut -c1-400
