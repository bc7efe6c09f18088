#!/bin/bash
# r2m (1 GPU): K1 per-tap table, K1b slab kernel: parity tests, K1 / K1b against the HBM roofline, headline + net lines
TAG=${1:-r2m}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -6 | cut -c1-300
timeout 200 python tools/bench_k1.py > $O/${TAG}_bench_k1.txt 2>&1; cat $O/${TAG}_bench_k1.txt | grep -v Warn
REPMODE_K1B_LEGACY=1 K1_ONLY=512 timeout 100 python tools/bench_k1.py 2>&1 | grep -v Warn | sed 's/^/legacy K1b: /'
REPMODE_BENCH_FAST=1 timeout 200 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "headline: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench.json | head -1)"
timeout 400 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train.json 2> $O/${TAG}_net_train.err
echo "net_train: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_train.json | head -1)"
echo done
