#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/c7_pytest.log
BAM3D_LIB=build/variants/lib_experiments.so timeout -s KILL 200 python -m pytest tests/test_narrow_gpu.py -m gpu -x -q -k "execution_options or lane_refill" 2>&1 | tail -5 | tee gpurun_out/c7_pytest_experiments.log
( time timeout -s KILL 500 python bench.py --steps 10 --warmup 3 ) > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/c7_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/c7_bench.json'))
    print('c5', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k,v in d['extra'].items():
        print(k, v['value'], v['ms_per_step'], 'e2e', v['e2e']['value'])
    print(json.dumps(d['extra']['c4']['extra'])[:3000])
except Exception as e:
    print('bench parse failed', e)
PY
