#!/bin/bash
# r2i2 (2 GPUs): the D-sharded block with the two-phase K4 -- parity over peer memory and the headline line at N = 2.
O=gpurun_out; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1 REPMODE_NO_BUILD=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29542 tests/check_sharded_block.py --comm peer > $O/r2i2_shard_peer_n2.log 2>&1
echo "sharded block (peer, 2 GPUs) exit $?"; grep -E "SHARDED_BLOCK_OK|FAILED|Error|^\[peer" $O/r2i2_shard_peer_n2.log | tail -4 | cut -c1-300
timeout 300 $TR --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2i2_bench_n2.json 2> $O/r2i2_bench_n2.err
echo "bench N=2 exit $?"; grep -o '"ms_per_step": [0-9.]*\|"nccl_ms_per_step": [0-9.]*\|"replicas_ms_per_step": [0-9.]*' $O/r2i2_bench_n2.json | head -6 | tr '\n' ' '; echo
tail -2 $O/r2i2_bench_n2.err | cut -c1-300
echo done
