#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 400 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tests/tools/gather_2gpu.py 2>&1 | grep -E "gather ok|Error|error|assert" | head -5
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_c5_8gpu_r2.json 2> gpurun_out/c25_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_c5_8gpu_r2.json').read().strip().splitlines()[-1])
    print('N=8 c5 %.4g pairs/s %.2f ms/step' % (d['value'], d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'], 'contact_all %.4g' % d['contact_all']['value'], 'intersect_all %.4g' % d['intersect_all']['value'])
    for k, v in d['extra'].items():
        print(' ', k, '%.4g (%.3f ms) e2e %.4g' % (v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value', 0)), v.get('scaling'))
except Exception as e:
    print('parse failed', e)
PY
tail -3 gpurun_out/c25_bench8.err | cut -c1-300
