#!/bin/bash
# D-sharded headline bench at N GPUs (+ sharded-block parity over peer memory at N ranks):  bash tools/gpu_scale.sh <tag> <N>
TAG=${1:-r2s}; N=${2:-4}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    tests/check_sharded_block.py --comm peer > $O/${TAG}_shard_peer_n$N.log 2>&1
echo "sharded block (peer, $N GPUs) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|^\[peer" $O/${TAG}_shard_peer_n$N.log | tail -4 | cut -c1-250
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
echo "bench N=$N exit $?"; TAGN=${TAG}_bench_n$N python - <<'PY'
import json, os
d=json.loads(open('gpurun_out/%s.json' % os.environ['TAGN']).read().strip().splitlines()[-1])
for k in ['n_gpus','value','ms_per_step','nccl_ms_per_step','replicas_ms_per_step','exchange_step_us']:
    print(k, d.get(k))
print('e2e ms', d['e2e']['ms_per_step'], 'sustained ms', d['sustained']['ms_per_step'], d['sustained']['clocks'])
PY
tail -3 $O/${TAG}_bench_n$N.err | cut -c1-300
echo done
