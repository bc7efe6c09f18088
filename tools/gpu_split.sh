#!/bin/bash
# A/B visit for the split-tap wgrad (REPMODE_WGRAD_SPLIT=1): whole GPU suite with it as the default tcgen05 wgrad, bench, ncu.
TAG=${1:-r1g}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_WGRAD_SPLIT=1
timeout 100 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 60 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -8 | cut -c1-200
REPMODE_BENCH_FAST=1 timeout 60 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_split.json 2> $O/${TAG}_bench.err
grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_split.json | head -1; grep -o '"wgrad_ms": [0-9.]*' $O/${TAG}_bench_split.json
REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 60 ncu --set full --clock-control none --import-source on \
  -k regex:'wgrad_split_kernel' -s 3 -c 1 -o $O/${TAG}_full_wgrad -f python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_full_wgrad.log 2>&1
[ -s $O/${TAG}_full_wgrad.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_wgrad.ncu-rep --page raw --csv > $O/${TAG}_full_wgrad_raw.csv 2>/dev/null
ls -la $O/${TAG}_full_wgrad.ncu-rep 2>&1 | cut -c20-
echo done
