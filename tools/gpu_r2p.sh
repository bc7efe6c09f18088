#!/bin/bash
# r2p (1 GPU): predict blend kernels, fused Adam, TF32 stride-2 GEMMs in training, split-K: tests + net lines
TAG=${1:-r2p}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
timeout 300 python tools/profile_net.py --train --batch 4 > $O/${TAG}_profile_net_train4.txt 2>&1; grep -E "step, batch|sum of MoDEConv|kernel time| us x" $O/${TAG}_profile_net_train4.txt | head -24 | cut -c1-150
timeout 300 python bench.py --config net_fwd --steps 20 --warmup 5 > $O/${TAG}_net_fwd.json 2> $O/${TAG}_net_fwd.err
echo "net_fwd: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_fwd.json | head -1)"
timeout 400 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train.json 2> $O/${TAG}_net_train.err
echo "net_train: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_train.json | head -1) launches $(grep -o '"gpu_launches": [0-9]*' $O/${TAG}_net_train.json)"; tail -3 $O/${TAG}_net_train.err
echo done
