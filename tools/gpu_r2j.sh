#!/bin/bash
# r2j (1 GPU): pair kernel for K > 32 (chunk passes), eval path fixes; per-layer profile of the U-Net (eval and train)
TAG=${1:-r2j}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
timeout 300 python tools/profile_net.py > $O/${TAG}_profile_net_eval.txt 2>&1; cat $O/${TAG}_profile_net_eval.txt | grep -v Warn | cut -c1-150
REPMODE_PAIR_MAXK=32 timeout 300 python tools/profile_net.py > $O/${TAG}_profile_net_eval_maxk32.txt 2>&1; grep -E "step, batch|sum of MoDEConv|kernel time|->" $O/${TAG}_profile_net_eval_maxk32.txt | cut -c1-150
timeout 300 python tools/profile_net.py --train --batch 1 > $O/${TAG}_profile_net_train.txt 2>&1; cat $O/${TAG}_profile_net_train.txt | grep -v Warn | cut -c1-150
timeout 300 python bench.py --config net_fwd --steps 20 --warmup 5 > $O/${TAG}_net_fwd.json 2> $O/${TAG}_net_fwd.err
echo "net_fwd: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_fwd.json | head -1)"
timeout 400 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train.json 2> $O/${TAG}_net_train.err
echo "net_train: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_train.json | head -1)"
echo done
