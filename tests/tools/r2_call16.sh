#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/c16_$name.json 2> gpurun_out/c16_$name.err
  python - $name <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/c16_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'c5 %.4g pairs/s %.2f ms/step' % (d['value'], d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'], 'contact_all %.4g' % d['contact_all']['value'], 'intersect_all %.4g' % d['intersect_all']['value'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/c16_n1.json 2> gpurun_out/c16_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c16_n1.json').read().strip().splitlines()[-1])
print('n1', 'c5 %.4g pairs/s %.2f ms/step' % (d['value'], d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'], 'contact_all %.4g' % d['contact_all']['value'])
PY
run prio0 BAM3D_COMM_PRIORITY=0
run prio1 BAM3D_COMM_PRIORITY=1
run prio1_cta8 BAM3D_COMM_PRIORITY=1 NCCL_MAX_CTAS=8
run prio1_cta4 BAM3D_COMM_PRIORITY=1 NCCL_MAX_CTAS=4
run prio0_cta4 BAM3D_COMM_PRIORITY=0 NCCL_MAX_CTAS=4
