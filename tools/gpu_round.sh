#!/bin/bash
# One GPU-box visit: bench, parity tests, per-kernel breakdown, ncu launch list + full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [skip_tests]
# Everything lands in gpurun_out/<tag>_*; tools/ncu_summarise.py turns the ncu CSVs into profiles/ summaries.
TAG=${1:-r1}
SKIP_TESTS=${2:-0}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1

echo "== bench (ours)"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -c 600 $O/${TAG}_bench.err
cat $O/${TAG}_bench.json | cut -c1-1500

if [ "$SKIP_TESTS" != "1" ]; then
  echo "== pytest -m gpu"
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1
  echo "pytest exit $?"
  tail -5 $O/${TAG}_pytest.log
fi

echo "== step breakdown"
timeout 300 python tools/step_breakdown.py > $O/${TAG}_breakdown.log 2>&1
cp $O/step_breakdown.json $O/${TAG}_step_breakdown.json 2>/dev/null
tail -3 $O/${TAG}_breakdown.log | cut -c1-600

echo "== conv issue-warp profile"
timeout 300 python tools/profile_conv.py 3 4 > $O/${TAG}_conv_waits.log 2>&1
cut -c1-700 $O/${TAG}_conv_waits.log | tail -4

echo "== ncu launch list (steps only)"
REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_list.log 2>&1
wc -l $O/${TAG}_launches.csv

echo "== ncu full: conv (fwd + dgrad), wgrad with source; the streaming kernels without"
ncu_full() {   # name, kernel regex, skip, count, extra flags
  REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 600 ncu --set full --clock-control none $5 \
    -k regex:"$2" -s $3 -c $4 -o $O/${TAG}_full_$1 -f python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_full_$1.log 2>&1
  ncu -i $O/${TAG}_full_$1.ncu-rep --page raw --csv > $O/${TAG}_full_$1_raw.csv 2>/dev/null
  ls -la $O/${TAG}_full_$1.ncu-rep | cut -c20-
}
ncu_full conv 'conv3d_pair|conv3d_umma' 6 2 "--import-source on"
ncu_full wgrad 'wgrad_umma' 3 1 "--import-source on"
ncu_full stream 'bn_|reparam|cast_f16|pack_dgrad|gate_bwd|wgrad_reduce' 33 11 ""

echo "== reference arm"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>/dev/null
cut -c1-400 $O/${TAG}_bench_ref.json
echo done
