#!/bin/bash
# r2z (N GPUs): BASELINE.json config 4 (N = 4) / config 5 (N = 8) bench line, the D-sharded headline line at N, sharded-block
# parity over peer memory at N ranks, whole-Net sharded parity (N <= 4).
#   bash tools/gpu_r2z.sh <tag> <N>
TAG=${1:-r2z}; N=${2:-8}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
CFG=cfg4; [ "$N" = "8" ] && CFG=cfg5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
echo "bench N=$N exit $?"; TAGN=${TAG}_bench_n$N python - <<'PY'
import json, os
d=json.loads(open('gpurun_out/%s.json' % os.environ['TAGN']).read().strip().splitlines()[-1])
for k in ['n_gpus','value','ms_per_step','nccl_ms_per_step','replicas_ms_per_step','exchange_step_us']:
    print(k, d.get(k))
print('e2e ms', d['e2e']['ms_per_step'], 'sustained ms', d['sustained']['ms_per_step'], d['sustained']['clocks'])
PY
tail -3 $O/${TAG}_bench_n$N.err | cut -c1-300
timeout 500 $TR --master-port 29546 bench.py --gpus $N --config $CFG --steps 5 --warmup 3 > $O/${TAG}_${CFG}_n$N.json 2> $O/${TAG}_${CFG}_n$N.err
echo "$CFG N=$N exit $?"; grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.e+]*\|"frac": [0-9.]*\|"mem_gb": [0-9.]*' $O/${TAG}_${CFG}_n$N.json | head -8 | tr '\n' ' '; echo
tail -4 $O/${TAG}_${CFG}_n$N.err | cut -c1-300
timeout 200 $TR --master-port 29542 tests/check_sharded_block.py --comm peer > $O/${TAG}_shard_peer_n$N.log 2>&1
echo "sharded block (peer, $N GPUs) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|^\[peer" $O/${TAG}_shard_peer_n$N.log | tail -4 | cut -c1-300
if [ "$N" -le 4 ]; then
timeout 400 $TR --master-port 29533 tests/check_sharded.py > $O/${TAG}_check_sharded_n$N.log 2>&1
echo "check_sharded (stage + whole Net vs oracle, $N GPUs) exit $?"; grep -E "SHARDED_CHECK_OK|FAILED|Error|^\[" $O/${TAG}_check_sharded_n$N.log | tail -12 | cut -c1-300
fi
echo done
