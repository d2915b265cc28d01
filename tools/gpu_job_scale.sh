#!/bin/bash
# usage: gpu_job_scale.sh <N> [config] [steps]  -- bench.py under torchrun on N GPUs of one box (as the driver launches it)
N=${1:-2}; CFG=${2:-C2}; STEPS=${3:-10}
mkdir -p gpurun_out
TAG=$(echo $CFG | tr 'A-Z' 'a-z')
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $CFG --steps $STEPS --warmup 3 --no-cpu > gpurun_out/r2u_bench_${TAG}_n$N.json 2> gpurun_out/r2u_bench_${TAG}_n$N.err
echo "rc=$?"; tail -c 400 gpurun_out/r2u_bench_${TAG}_n$N.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/r2u_bench_${TAG}_n$N.json').read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'] and (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['breakdown_ms']), 'hdw_sharded', d['detail'].get('hdw_sharded_over_ranks'))"
