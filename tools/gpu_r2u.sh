#!/bin/bash
# r2u (1 GPU): validate the K1b register kernel / 16-byte dgrad pack / gate-backward chunks / stem cast+pad:
# parity tests, K1 table, net_train + per-kernel profile, the headline bench line.
TAG=${1:-r2u}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
timeout 200 python tools/bench_k1.py > $O/${TAG}_bench_k1.txt 2>&1; grep -v Warn $O/${TAG}_bench_k1.txt | tail -12
timeout 400 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train.json 2> $O/${TAG}_net_train.err
echo "net_train: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_train.json | head -1) launches $(grep -o '"gpu_launches": [0-9]*' $O/${TAG}_net_train.json)"; tail -3 $O/${TAG}_net_train.err
timeout 300 python tools/profile_net.py --train --batch 4 > $O/${TAG}_profile_net_train4.txt 2>&1
grep -v Warn $O/${TAG}_profile_net_train4.txt | sed -n '1p;21,45p' | cut -c1-150
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench.json | head -3 | tr '\n' ' '; tail -3 $O/${TAG}_bench.err | cut -c1-300
echo done
