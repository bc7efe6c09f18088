#!/bin/bash
# r2e (1 GPU): full GPU suite, sharded block over peer memory (2 ranks on one GPU), bench (+K1 rows kernel, single-wave grids)
TAG=${1:-r2e}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
for comm in peer; do
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    tests/check_sharded_block.py --comm $comm --same-device > $O/${TAG}_shard_$comm.log 2>&1
  echo "sharded block ($comm, 2 ranks on one GPU) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|error" $O/${TAG}_shard_$comm.log | tail -4 | cut -c1-300
done
timeout 200 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cut -c1-400 $O/${TAG}_bench.json; echo; grep -o '"others": {[^}]*}' $O/${TAG}_bench.json; grep -o '"sustained": {[^}]*}' $O/${TAG}_bench.json | cut -c1-300; grep -o '"e2e": {[^}]*}' $O/${TAG}_bench.json | cut -c1-200; tail -3 $O/${TAG}_bench.err
python tools/step_breakdown.py > $O/${TAG}_breakdown.log 2>&1; tail -22 $O/${TAG}_breakdown.log | cut -c1-120
echo done
