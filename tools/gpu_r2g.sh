#!/bin/bash
# r2g (2 GPUs): does the D-sharded whole-Net train step capture as a CUDA graph with its NCCL collectives inside?
# cfg4's per-rank slab shape (16 planes of 256x256 per rank) on 2 ranks through the dry-run hook, graph vs eager; plus smoke().
O=gpurun_out; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1 REPMODE_NO_BUILD=1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/r2g_smoke.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
REPMODE_BENCH_CFG_DIMS=32,256,256 timeout 150 $TR --master-port 29551 bench.py --gpus 2 --config cfg4 --steps 5 --warmup 3 > $O/r2g_cfg4dims_graph.json 2> $O/r2g_cfg4dims_graph.err
echo "graph exit $?"; grep -o '"ms_per_step": [0-9.]*\|"notes": \[[^]]*\]' $O/r2g_cfg4dims_graph.json | head -4; tail -2 $O/r2g_cfg4dims_graph.err | cut -c1-300
REPMODE_BENCH_GRAPH_SHARDED=0 REPMODE_BENCH_CFG_DIMS=32,256,256 timeout 150 $TR --master-port 29552 bench.py --gpus 2 --config cfg4 --steps 5 --warmup 3 > $O/r2g_cfg4dims_eager.json 2> $O/r2g_cfg4dims_eager.err
echo "eager exit $?"; grep -o '"ms_per_step": [0-9.]*\|"notes": \[[^]]*\]' $O/r2g_cfg4dims_eager.json | head -4
echo done
