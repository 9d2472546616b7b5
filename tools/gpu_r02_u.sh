#!/bin/bash
# Round 2, GPU session U (2 x B200): the sharded default job with the optional gather of the output slabs on rank 0.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 3 --warmup 3 --gather --no-cpu --extras none > $O/u_bench_2gpu_gather.json 2> $O/u_bench_2gpu_gather.err
echo "2gpu gather rc=$?"
python -c "
import json
d=json.load(open('$O/u_bench_2gpu_gather.json'))
print('value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'verified',d['verified'])
print('gather',d.get('gather'))
"
tail -5 $O/u_bench_2gpu_gather.err
