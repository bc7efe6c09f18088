#!/bin/bash
# r2i (1 GPU): eval path (BN folded into K2, fp16 activations, Net graph), frozen-BN backward, sharded formulation cost,
# whole-Net config lines
TAG=${1:-r2i}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
REPMODE_BENCH_FAST=1 timeout 200 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "headline: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench.json | head -1)"
REPMODE_BENCH_SHARDED_LOCAL=1 REPMODE_BENCH_FAST=1 timeout 200 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_shl.json 2> $O/${TAG}_bench_shl.err
echo "sharded formulation, no neighbours: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_shl.json | head -1)"; tail -2 $O/${TAG}_bench_shl.err | cut -c1-300
REPMODE_BENCH_SHARDED_LOCAL=1 python tools/step_breakdown.py > $O/${TAG}_breakdown_shl.log 2>&1; tail -22 $O/${TAG}_breakdown_shl.log | cut -c1-120 | head -16
timeout 300 python bench.py --config net_fwd --steps 20 --warmup 5 > $O/${TAG}_net_fwd.json 2> $O/${TAG}_net_fwd.err
echo "net_fwd exit $?"; cut -c1-700 $O/${TAG}_net_fwd.json; tail -3 $O/${TAG}_net_fwd.err | cut -c1-300
timeout 400 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train.json 2> $O/${TAG}_net_train.err
echo "net_train exit $?"; cut -c1-900 $O/${TAG}_net_train.json; tail -3 $O/${TAG}_net_train.err | cut -c1-300
echo done
