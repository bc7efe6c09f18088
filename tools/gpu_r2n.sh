#!/bin/bash
# r2n (1 GPU): split-K single-CTA conv for the deep small-volume levels: parity tests, per-layer profile of the U-Net (eval,
# train batch 1 and 4), net_fwd / net_train / headline lines
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
timeout 300 python tools/profile_net.py > $O/${TAG}_profile_net_eval.txt 2>&1; cat $O/${TAG}_profile_net_eval.txt | grep -v Warn | cut -c1-150
REPMODE_UMMA_SPLITK=1 timeout 300 python tools/profile_net.py > $O/${TAG}_profile_net_eval_nosplit.txt 2>&1; grep -E "step, batch|sum of MoDEConv|kernel time" $O/${TAG}_profile_net_eval_nosplit.txt | cut -c1-150
timeout 300 python tools/profile_net.py --train --batch 1 > $O/${TAG}_profile_net_train.txt 2>&1; cat $O/${TAG}_profile_net_train.txt | grep -v Warn | cut -c1-150
timeout 300 python tools/profile_net.py --train --batch 4 > $O/${TAG}_profile_net_train4.txt 2>&1; cat $O/${TAG}_profile_net_train4.txt | grep -v Warn | cut -c1-150
timeout 300 python bench.py --config net_fwd --steps 20 --warmup 5 > $O/${TAG}_net_fwd.json 2> $O/${TAG}_net_fwd.err
echo "net_fwd: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_fwd.json | head -1)"
timeout 400 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train.json 2> $O/${TAG}_net_train.err
echo "net_train: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_train.json | head -1) launches $(grep -o '"gpu_launches": [0-9]*' $O/${TAG}_net_train.json)"
REPMODE_BENCH_FAST=1 timeout 200 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "headline: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench.json | head -1)"
echo done
