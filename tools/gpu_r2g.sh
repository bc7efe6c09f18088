#!/bin/bash
# r2g (1 GPU): BN streaming loads / trivial-planes fast path: tests + bench + per-kernel breakdown
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
REPMODE_BENCH_FAST=1 timeout 200 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench.json | head -1
python tools/step_breakdown.py > $O/${TAG}_breakdown.log 2>&1; tail -22 $O/${TAG}_breakdown.log | cut -c1-120 | head -12
echo done
