#!/bin/bash
# r2d (1 GPU): full GPU suite, sharded block (2 ranks on one GPU), bench, ncu full of the BatchNorm-backward passes
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
for comm in gloo peer; do
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    tests/check_sharded_block.py --comm $comm --same-device > $O/${TAG}_shard_$comm.log 2>&1
  echo "sharded block ($comm, 2 ranks on one GPU) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|error" $O/${TAG}_shard_$comm.log | tail -4 | cut -c1-300
done
timeout 200 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cut -c1-1800 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 120 ncu --set full --clock-control none \
  --import-source on -k regex:'bn_bwd_reduce_vec_kernel|bn_bwd_apply_vec_kernel|cast_f16_kernel|bn_apply_kernel' -s 12 -c 4 -o $O/${TAG}_full_bn -f python bench.py --steps 3 --warmup 3 \
  > $O/${TAG}_ncu_full_bn.log 2>&1
[ -s $O/${TAG}_full_bn.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_bn.ncu-rep --page raw --csv > $O/${TAG}_full_bn_raw.csv 2>/dev/null
[ -s $O/${TAG}_full_bn.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_bn.ncu-rep --page source --csv > $O/${TAG}_full_bn_src.csv 2>/dev/null
ls -la $O/${TAG}_full_bn* | cut -c20-
echo done
