#!/bin/bash
# why does ncclBroadcast of the class matrix take 8 ms on 4 of the box's 8 GPUs (0.7 ms on 2 and on 8)?
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --config C2 --steps 3 --warmup 3 --no-cpu > gpurun_out/n4_$tag.json 2> gpurun_out/n4_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/n4_$tag.json').read().strip().splitlines()[-1]); print('$tag', d['e2e']['ms_per_step'], d['e2e']['breakdown_ms']['load_codes (H2D on rank 0 + ncclBroadcast)'])"; }
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING
grep -i "broadcast\|Algo\|channels\|NVLS\|via" gpurun_out/n4_default.err | head -30 | cut -c1-200
run ring NCCL_ALGO=Ring
run simple NCCL_PROTO=Simple
