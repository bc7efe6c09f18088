#!/bin/bash
# BASELINE.json configs 4 / 5 (whole U-Net train step on ONE volume, D-sharded) at N GPUs, plus the sharded-block parity
# check and the headline D-sharded bench line:   bash tools/gpu_cfg.sh <tag> <N> [dry D,H,W]
#   N = 4 -> cfg4 (64x256x256), N = 8 -> cfg5 (128x512x512); "dry 32,128,128" runs cfg4's code path on a small volume at any N
TAG=${1:-r2c}; N=${2:-4}; MODE=${3:-full}; DIMS=${4:-32,128,128}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
CFG=cfg4; [ "$N" = "8" ] && CFG=cfg5
if [ "$MODE" = "dry" ]; then export REPMODE_BENCH_CFG_DIMS=$DIMS; fi
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    tests/check_sharded_block.py --comm peer > $O/${TAG}_shard_peer_n$N.log 2>&1
echo "sharded block (peer, $N GPUs) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|^\[peer" $O/${TAG}_shard_peer_n$N.log | tail -4 | cut -c1-250
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 \
  bench.py --gpus $N --config $CFG --steps 5 --warmup 3 > $O/${TAG}_${CFG}_n$N.json 2> $O/${TAG}_${CFG}_n$N.err
echo "$CFG N=$N exit $?"; grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.e+]*\|"frac": [0-9.]*\|"mem_gb": [0-9.]*' $O/${TAG}_${CFG}_n$N.json | head -8 | tr '\n' ' '; echo
tail -4 $O/${TAG}_${CFG}_n$N.err | cut -c1-300
unset REPMODE_BENCH_CFG_DIMS
if [ "$MODE" != "dry" ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
echo "bench N=$N exit $?"; TAGN=${TAG}_bench_n$N python - <<'PY'
import json, os
d=json.loads(open('gpurun_out/%s.json' % os.environ['TAGN']).read().strip().splitlines()[-1])
for k in ['n_gpus','value','ms_per_step','nccl_ms_per_step','replicas_ms_per_step','exchange_step_us']:
    print(k, d.get(k))
print('e2e ms', d['e2e']['ms_per_step'], 'sustained ms', d['sustained']['ms_per_step'], d['sustained']['clocks'])
PY
tail -3 $O/${TAG}_bench_n$N.err | cut -c1-300
fi
echo done
