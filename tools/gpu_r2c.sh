#!/bin/bash
# r2c: _ex conv / wgrad entry points, elementwise rewrites (BN passes, cast), D-sharded block on ONE GPU (2 ranks, gloo + peer)
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 -x > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
for comm in gloo peer; do
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    tests/check_sharded_block.py --comm $comm --same-device > $O/${TAG}_shard_$comm.log 2>&1
  echo "sharded block ($comm, 2 ranks on one GPU) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|error|collectives" $O/${TAG}_shard_$comm.log | tail -8 | cut -c1-400
done
ab() {   # name, env assignments...
  local name=$1; shift
  env "$@" REPMODE_BENCH_FAST=1 timeout 60 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
  echo "$name: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_$name.json | head -1) $(grep -o '"wgrad_ms": [0-9.]*' $O/${TAG}_bench_$name.json) $(grep -o '"conv_fwd_ms": [0-9.]*' $O/${TAG}_bench_$name.json)"
}
ab default REPMODE_NOOP=1
ab nokeep REPMODE_BN_L2_KEEP_MB=0
ab keep60 REPMODE_BN_L2_KEEP_MB=60
python tools/step_breakdown.py > $O/${TAG}_breakdown.log 2>&1; tail -25 $O/${TAG}_breakdown.log
echo done
