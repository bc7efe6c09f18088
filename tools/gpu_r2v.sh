#!/bin/bash
# r2v (N GPUs, default 2): early x-halo signal + dy-halo wait in front of K3 + block-form gate backward:
# sharded-block parity over peer memory, whole-Net sharded parity vs oracle, the D-sharded headline bench line, gate test.
TAG=${1:-r2v}; N=${2:-2}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -p no:cacheprovider --timeout 200 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -6 | cut -c1-300
timeout 200 $TR --master-port 29542 tests/check_sharded_block.py --comm peer > $O/${TAG}_shard_peer_n$N.log 2>&1
echo "sharded block (peer, $N GPUs) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|^\[peer" $O/${TAG}_shard_peer_n$N.log | tail -4 | cut -c1-400
timeout 400 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
echo "bench N=$N exit $?"; TAGN=${TAG}_bench_n$N python - <<'PY'
import json, os
d=json.loads(open('gpurun_out/%s.json' % os.environ['TAGN']).read().strip().splitlines()[-1])
for k in ['n_gpus','value','ms_per_step','nccl_ms_per_step','replicas_ms_per_step','exchange_step_us']:
    print(k, d.get(k))
print('e2e ms', d['e2e']['ms_per_step'], 'sustained ms', d['sustained']['ms_per_step'], d['sustained']['clocks'])
PY
tail -3 $O/${TAG}_bench_n$N.err | cut -c1-300
timeout 400 $TR --master-port 29533 tests/check_sharded.py > $O/${TAG}_check_sharded_n$N.log 2>&1
echo "check_sharded (stage + whole Net vs oracle, $N GPUs) exit $?"; grep -E "SHARDED_CHECK_OK|FAILED|Error|^\[" $O/${TAG}_check_sharded_n$N.log | tail -12 | cut -c1-420
echo done
