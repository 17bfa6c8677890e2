mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r02i_pytest.log; tail -5 gpurun_out/r02i_pytest.log
(timeout 400 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02i_bench.err | tail -1) > gpurun_out/r02i_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r02i_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['library_baseline']['ms_per_step'] if d.get('library_baseline') else None, d['inversion_step'], {k:(v['ms_per_step'],v['launches']) for k,v in d['kernels'].items()})"
