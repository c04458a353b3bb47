#!/bin/bash
mkdir -p gpurun_out
show() { python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/c20_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'c5 %.4g pairs/s %.2f ms/step' % (d['value'], d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'], 'contact_all %.4g' % d['contact_all']['value'], 'intersect_all %.4g' % d['intersect_all']['value'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
run2() { name=$1; shift
  env "$@" timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra > gpurun_out/c20_$name.json 2> gpurun_out/c20_$name.err; show $name; grep "bam3d comm\|rror" gpurun_out/c20_$name.err | cut -c1-250 | head -4; }
run2 cta4 BAM3D_COMM_TRACE=1
run2 cta2 BAM3D_COMM_TRACE=1 BAM3D_NCCL_MAX_CTAS=2
run2 cta4b A=1
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/tools/gather_2gpu.py 2>&1 | grep -E "gather ok|Error|error|assert" | head -5
