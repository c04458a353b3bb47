#!/bin/bash
# 2 GPUs: the library's NCCL gather (device pointers, host buffers, frame queue) and the default bench under torchrun
mkdir -p gpurun_out
nvidia-smi -L
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/gather_2gpu.py > gpurun_out/c12_gather.log 2>&1; grep -E 'gather ok|Error|error|assert' gpurun_out/c12_gather.log | head -12
( time timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/c12_bench2.json 2> gpurun_out/c12_bench2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/c12_bench2.err
python - <<'PY'
import json
try:
    lines=[l for l in open('gpurun_out/c12_bench2.json') if l.startswith('{')]
    d=json.loads(lines[-1])
    print('c5', d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'contact_all', d['contact_all']['value'], 'intersect_all', d['intersect_all']['value'])
    for k,v in d['extra'].items():
        print(k, v['value'], v['ms_per_step'], 'e2e', v['e2e']['value'])
except Exception as e:
    print('bench parse failed', e)
PY
( time timeout -s KILL 300 python bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) 2>&1 | tail -c 600
