mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r02l_pytest.log; tail -3 gpurun_out/r02l_pytest.log
(timeout 400 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02l_bench.err | tail -1) > gpurun_out/r02l_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r02l_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], {k:(v['ms_per_step'],v['launches']) for k,v in d['kernels'].items()})"
bash tools/gpu_profile_r02.sh > gpurun_out/r02l_profile.log 2>&1; tail -12 gpurun_out/r02l_profile.log
