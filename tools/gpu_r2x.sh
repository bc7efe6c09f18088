#!/bin/bash
# r2x (1 GPU): step-level grouped K1 plan: parity tests, net_train with / without the plan, per-kernel profile.
TAG=${1:-r2x}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
timeout 400 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train.json 2> $O/${TAG}_net_train.err
echo "net_train (grouped K1): $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_train.json | head -1) launches $(grep -o '"gpu_launches": [0-9]*' $O/${TAG}_net_train.json)"; tail -2 $O/${TAG}_net_train.err | cut -c1-300
REPMODE_K1_GROUPED=0 timeout 400 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train_nogroup.json 2> $O/${TAG}_net_train_nogroup.err
echo "net_train (per-layer K1): $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_train_nogroup.json | head -1) launches $(grep -o '"gpu_launches": [0-9]*' $O/${TAG}_net_train_nogroup.json)"
timeout 300 python tools/profile_net.py --train --batch 4 > $O/${TAG}_profile_net_train4.txt 2>&1
grep -v Warn $O/${TAG}_profile_net_train4.txt | sed -n '1p;21,50p' | cut -c1-150
echo done
