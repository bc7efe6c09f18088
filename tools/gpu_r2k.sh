#!/bin/bash
# r2k (1 GPU): new parity tests (sharded on one GPU, full-width train gradients), K1 / K1b against the HBM roofline + ncu
TAG=${1:-r2k}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 -s > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|full-width train-step" $O/${TAG}_pytest.log | tail -12 | cut -c1-1200
timeout 200 python tools/bench_k1.py > $O/${TAG}_bench_k1.txt 2>&1; cat $O/${TAG}_bench_k1.txt | grep -v Warn
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'reparam_bwd_kernel|reparam_fwd_rows_kernel|pack_dgrad_kernel' -s 60 -c 3 \
  -o $O/${TAG}_full_k1 -f python tools/bench_k1.py > $O/${TAG}_ncu_k1.log 2>&1
[ -s $O/${TAG}_full_k1.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_k1.ncu-rep --page raw --csv > $O/${TAG}_full_k1_raw.csv 2>/dev/null
[ -s $O/${TAG}_full_k1.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_k1.ncu-rep --page source --csv > $O/${TAG}_full_k1_src.csv 2>/dev/null
ls -la $O/${TAG}_full_k1* | cut -c20-
echo done
