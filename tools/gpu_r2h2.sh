#!/bin/bash
# r2h2 (2 GPUs): exchange steps fused into the kernels: parity (peer, fused and unfused), bench N=2 fused vs unfused + profile
TAG=${1:-r2h2}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    tests/check_sharded_block.py --comm peer --steps 3 > $O/${TAG}_shard_peer.log 2>&1
echo "sharded block (peer fused, 2 GPUs) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|^\[peer" $O/${TAG}_shard_peer.log | tail -4 | cut -c1-300
REPMODE_FUSED_EXCHANGE=0 timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    tests/check_sharded_block.py --comm peer > $O/${TAG}_shard_peer_unfused.log 2>&1
echo "sharded block (peer unfused, 2 GPUs) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error" $O/${TAG}_shard_peer_unfused.log | tail -3 | cut -c1-300
REPMODE_BENCH_PROFILE=1 REPMODE_BENCH_FAST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/${TAG}_bench_n2_prof.json 2> $O/${TAG}_bench_n2_prof.err
echo "bench N=2 fused (profile) exit $?"; grep " us x" $O/${TAG}_bench_n2_prof.err | head -22 | cut -c1-130; grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_n2_prof.json | head -1
REPMODE_FUSED_EXCHANGE=0 REPMODE_BENCH_FAST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/${TAG}_bench_n2_unfused.json 2> $O/${TAG}_bench_n2_unfused.err
echo "bench N=2 unfused exit $?"; grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_n2_unfused.json | head -1
timeout 120 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 -k "sharded or parity or net" > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -5 | cut -c1-300
echo done
