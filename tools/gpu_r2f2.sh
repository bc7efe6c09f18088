#!/bin/bash
# r2f2 (2 GPUs): D-sharded headline bench at N = 2 after the exchange trims, per-kernel profile of the sharded step
TAG=${1:-r2f2}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    tests/check_sharded_block.py --comm peer > $O/${TAG}_shard_peer.log 2>&1
echo "sharded block (peer, 2 GPUs) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error" $O/${TAG}_shard_peer.log | tail -3 | cut -c1-300
REPMODE_BENCH_PROFILE=1 REPMODE_BENCH_FAST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/${TAG}_bench_n2_prof.json 2> $O/${TAG}_bench_n2_prof.err
echo "bench N=2 (profile) exit $?"; grep " us x" $O/${TAG}_bench_n2_prof.err | head -30; grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_n2_prof.json | head -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
echo "bench N=2 exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f2_bench_n2.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','nccl_ms_per_step','replicas_ms_per_step','exchange_step_us']:
    print(k, d.get(k))
print(d['e2e']['ms_per_step'], d['sustained']['ms_per_step'])
PY
./tools/stream_probe > $O/${TAG}_stream_probe.txt 2>&1; cat $O/${TAG}_stream_probe.txt
echo done
