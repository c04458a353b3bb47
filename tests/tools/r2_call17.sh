#!/bin/bash
mkdir -p gpurun_out
show() { python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/c17_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'c5 %.4g pairs/s %.2f ms/step' % (d['value'], d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'], 'contact_all %.4g' % d['contact_all']['value'], 'intersect_all %.4g' % d['intersect_all']['value'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
run2() { name=$1; shift
  env "$@" timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/c17_$name.json 2> gpurun_out/c17_$name.err; show $name; }
BAM3D_BENCH_SLICE=1 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/c17_n1s1.json 2> gpurun_out/c17_n1s1.err; show n1s1
CUDA_VISIBLE_DEVICES=1 BAM3D_BENCH_SLICE=0 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/c17_n1gpu1.json 2> gpurun_out/c17_n1gpu1.err; show n1gpu1
run2 noexch BAM3D_BENCH_NO_EXCHANGE=1
run2 default A=1
