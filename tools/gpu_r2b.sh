#!/bin/bash
# r2b: deep wgrad with one-add descriptors as the default: bit-exact tests, bench A/B against the split kernel, ncu capture
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 200 python -m pytest tests/test_gpu_umma.py -q -p no:cacheprovider --timeout 60 > $O/${TAG}_pytest_umma.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest_umma.log | tail -12 | cut -c1-200
ab() {   # name, env assignments...
  local name=$1; shift
  env "$@" REPMODE_BENCH_FAST=1 timeout 60 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
  echo "$name: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_$name.json | head -1) $(grep -o '"wgrad_ms": [0-9.]*' $O/${TAG}_bench_$name.json) $(grep -o '"conv_fwd_ms": [0-9.]*' $O/${TAG}_bench_$name.json)"
}
ab default REPMODE_NOOP=1
ab split REPMODE_WGRAD_SPLIT=1
REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 60 ncu --set full --clock-control none \
  --import-source on -k regex:'wgrad_deep_kernel' -s 3 -c 1 -o $O/${TAG}_full_wgrad_deep -f python bench.py --steps 3 --warmup 3 \
  > $O/${TAG}_ncu_full_wgrad_deep.log 2>&1
[ -s $O/${TAG}_full_wgrad_deep.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_wgrad_deep.ncu-rep --page raw --csv > $O/${TAG}_full_wgrad_deep_raw.csv 2>/dev/null
[ -s $O/${TAG}_full_wgrad_deep.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_wgrad_deep.ncu-rep --page source --csv > $O/${TAG}_full_wgrad_deep_src.csv 2>/dev/null
echo done
