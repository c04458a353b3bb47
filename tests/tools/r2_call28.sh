#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py > gpurun_out/bench_c5_r2_final.json 2> gpurun_out/c28_bench.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/bench_c5_reference_r2_final.json 2> gpurun_out/c28_ref.err; echo "ref rc=$?"
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/bench_c5_2gpu_noextra_r2_final.json 2> gpurun_out/c28_b2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c5_r2_final.json').read().strip().splitlines()[-1])
print('c5 %.4g %.2f ms' % (d['value'], d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'], 'contact_all %.4g' % d['contact_all']['value'], 'intersect_all %.4g' % d['intersect_all']['value'])
print('cpu', {k: d['cpu_baseline'][k] for k in ('value', 'wall_value', 'cores')})
for k, v in d['extra'].items():
    print(' ', k, '%.4g (%.3f ms) e2e %.4g' % (v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value', 0)), 'cpu %.4g' % v['cpu_baseline']['value'])
r = json.loads(open('gpurun_out/bench_c5_reference_r2_final.json').read().strip().splitlines()[-1])
print('reference arm %.4g wall %.4g' % (r['value'], r['cpu_baseline']['wall_value']), {k: '%.4g' % v['value'] for k, v in r.get('extra', {}).items()})
d2 = json.loads(open('gpurun_out/bench_c5_2gpu_noextra_r2_final.json').read().strip().splitlines()[-1])
print('N=2 c5 %.4g %.2f ms' % (d2['value'], d2['ms_per_step']), 'e2e %.4g' % d2['e2e']['value'], 'contact_all %.4g' % d2['contact_all']['value'])
PY
