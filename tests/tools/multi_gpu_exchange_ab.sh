#!/bin/bash
# A/B of the contact exchange on N GPUs (default 8): the multi-rank check, then the C5 line with the all-gather payload, grouped
# broadcasts, 8 thread blocks and no exchange (BAM3D_COMM_TRACE prints what the exchange costs). Run with gpurun --gpus N.
mkdir -p gpurun_out
N=${1:-8}
show() { python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/xab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'c5 %.4g pairs/s %.2f ms/step' % (d['value'], d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'], 'contact_all %.4g' % d['contact_all']['value'], 'intersect_all %.4g' % d['intersect_all']['value'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
runN() { name=$1; shift
  env "$@" timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/xab_$name.json 2> gpurun_out/xab_$name.err; show $name; grep "bam3d comm rank [07]\]\|rror" gpurun_out/xab_$name.err | cut -c1-260 | head -4; }
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/tools/gather_2gpu.py 2>&1 | grep -E "gather ok|Error|error|assert" | head -5
runN allgather BAM3D_COMM_TRACE=1
runN bcast4 BAM3D_COMM_TRACE=1 BAM3D_COMM_ALLGATHER=0
runN allgather8 BAM3D_COMM_TRACE=1 BAM3D_NCCL_MAX_CTAS=8
runN noexch BAM3D_BENCH_NO_EXCHANGE=1
