#!/bin/bash
# r2d2 (2 GPUs): sharded block parity over NCCL and peer memory, then the D-sharded headline bench at N = 2
TAG=${1:-r2d2}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
for comm in nccl peer; do
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    tests/check_sharded_block.py --comm $comm > $O/${TAG}_shard_$comm.log 2>&1
  echo "sharded block ($comm, 2 GPUs) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|error|\[" $O/${TAG}_shard_$comm.log | tail -6 | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
echo "bench N=2 exit $?"; cut -c1-2500 $O/${TAG}_bench_n2.json; tail -5 $O/${TAG}_bench_n2.err | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
  tests/check_sharded.py > $O/${TAG}_shard_stage.log 2>&1
echo "sharded stage/net (NCCL, 2 GPUs) exit $?"; grep -E "SHARDED_CHECK_OK|FAILED|Error" $O/${TAG}_shard_stage.log | tail -4 | cut -c1-300
echo done
