#!/bin/bash
# First GPU visit of the next round: validate and A/B the experiments prepared at the end of round 1.
# Every step under a short timeout; ~2 minutes in total.   bash tools/gpu_next.sh [tag]
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
echo "== wgrad bit-exact tests incl. the experimental deep-tile kernel"
REPMODE_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_umma.py -k wgrad -q -p no:cacheprovider --timeout 60 > $O/${TAG}_pytest_wgrad.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest_wgrad.log | tail -12 | cut -c1-200
echo "== K1 wide-layer kernel: bit-identity test, then the whole-Net parity tests with it switched on"
REPMODE_TEST_EXPERIMENTAL=1 timeout 100 python -m pytest tests/test_gpu_parity.py -k "wide" -q -p no:cacheprovider --timeout 60 > $O/${TAG}_pytest_k1wide.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest_k1wide.log | tail -5 | cut -c1-200
REPMODE_K1_WIDE=1 timeout 120 python -m pytest tests/test_gpu_net.py -q -p no:cacheprovider --timeout 90 > $O/${TAG}_pytest_net_k1wide.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest_net_k1wide.log | tail -5 | cut -c1-200
ab() {   # name, env assignments...
  local name=$1; shift
  env "$@" REPMODE_BENCH_FAST=1 timeout 60 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
  echo "$name: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_$name.json | head -1) $(grep -o '"wgrad_ms": [0-9.]*' $O/${TAG}_bench_$name.json)"
}
ab default REPMODE_NOOP=1
ab wgrad_deep REPMODE_WGRAD_DEEP=1
ab k1_wide REPMODE_K1_WIDE=1
grep -o '"reparam_fwd_512x512_GBs": [0-9.]*' $O/${TAG}_bench_default.json $O/${TAG}_bench_k1_wide.json
ab bn_bps2 REPMODE_BN_REDUCE_BPS=2
ab bn_bps4 REPMODE_BN_REDUCE_BPS=4
echo "== ncu full of the deep-tile wgrad"
REPMODE_WGRAD_DEEP=1 REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 60 ncu --set full --clock-control none \
  --import-source on -k regex:'wgrad_deep_kernel' -s 3 -c 1 -o $O/${TAG}_full_wgrad_deep -f python bench.py --steps 3 --warmup 3 \
  > $O/${TAG}_ncu_full_wgrad_deep.log 2>&1
[ -s $O/${TAG}_full_wgrad_deep.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_wgrad_deep.ncu-rep --page raw --csv > $O/${TAG}_full_wgrad_deep_raw.csv 2>/dev/null
ls -la $O/${TAG}_full_wgrad_deep.ncu-rep 2>&1 | cut -c20-
echo done
