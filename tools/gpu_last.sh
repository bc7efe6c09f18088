#!/bin/bash
# Short confirmation visit: parity tests, bench, one ncu capture of K4 and the ncu launch list of the step.
TAG=${1:-r1h}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 100 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 60 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -8 | cut -c1-200
timeout 100 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cut -c1-330 $O/${TAG}_bench.json; echo; grep -o '"wgrad_ms": [0-9.]*' $O/${TAG}_bench.json
REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 60 ncu --set full --clock-control none --import-source on \
  -k regex:'wgrad_split_kernel|wgrad_umma_kernel' -s 3 -c 1 -o $O/${TAG}_full_wgrad -f python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_full_wgrad.log 2>&1
[ -s $O/${TAG}_full_wgrad.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_wgrad.ncu-rep --page raw --csv > $O/${TAG}_full_wgrad_raw.csv 2>/dev/null
ls -la $O/${TAG}_full_wgrad.ncu-rep 2>&1 | cut -c20-
REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none \
  -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_list.log 2>&1
wc -l $O/${TAG}_launches.csv
echo done
