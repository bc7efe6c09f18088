#!/bin/bash
# One GPU-box visit: parity tests, bench, A/B runs, ncu launch list + full captures -- every step under its own SHORT
# timeout, and the profiling steps only when tests and bench succeeded (a failing step must never eat the GPU budget).
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [tests|bench|ab|breakdown|ncu|ref ...]   (default: all but ab)
# Everything lands in gpurun_out/<tag>_*; tools/ncu_summarise.py turns the ncu CSVs into profiles/ summaries.
TAG=${1:-r1}
shift
STEPS=${*:-tests bench breakdown ncu ref}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
has() { case " $STEPS " in *" $1 "*) return 0;; *) return 1;; esac; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
OK=1

if has tests; then
  echo "== pytest -m gpu"
  timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 60 > $O/${TAG}_pytest.log 2>&1
  RC=$?
  echo "pytest exit $RC"
  grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -25 | cut -c1-300
  [ $RC -ne 0 ] && OK=0
fi

if has bench; then
  echo "== bench (ours)"
  timeout 200 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
  if [ ! -s $O/${TAG}_bench.json ]; then OK=0; echo "bench FAILED"; tail -c 1500 $O/${TAG}_bench.err; fi
  cut -c1-2600 $O/${TAG}_bench.json
fi

if has ab && [ $OK -eq 1 ]; then
  echo "== A/B (resident steps only): no stream overlap / split-tap wgrad"
  REPMODE_BENCH_FAST=1 REPMODE_OVERLAP=0 timeout 120 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_nooverlap.json 2>/dev/null
  cut -c1-330 $O/${TAG}_bench_nooverlap.json; echo
  if has split; then
    REPMODE_BENCH_FAST=1 REPMODE_WGRAD_SPLIT=1 timeout 120 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_split.json 2>/dev/null
    cut -c1-330 $O/${TAG}_bench_split.json; echo
    grep -o '"wgrad_ms": [0-9.]*' $O/${TAG}_bench.json $O/${TAG}_bench_split.json
  fi
fi

if has breakdown && [ $OK -eq 1 ]; then
  echo "== step breakdown"
  timeout 120 python tools/step_breakdown.py > $O/${TAG}_breakdown.log 2>&1
  cp $O/step_breakdown.json $O/${TAG}_step_breakdown.json 2>/dev/null
  tail -2 $O/${TAG}_breakdown.log | cut -c1-300
fi

if has ncu && [ $OK -eq 1 ]; then
  echo "== ncu launch list (steps only; kernels serialised, so the side-stream overlap is switched off)"
  REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none \
    -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_list.log 2>&1
  wc -l $O/${TAG}_launches.csv
  echo "== ncu full: conv (fwd + dgrad), wgrad with source; the streaming kernels without"
  ncu_full() {   # name, kernel regex, skip, count, extra flags
    REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 150 ncu --set full --clock-control none $5 \
      -k regex:"$2" -s $3 -c $4 -o $O/${TAG}_full_$1 -f python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_full_$1.log 2>&1
    [ -s $O/${TAG}_full_$1.ncu-rep ] && timeout 60 ncu -i $O/${TAG}_full_$1.ncu-rep --page raw --csv > $O/${TAG}_full_$1_raw.csv 2>/dev/null
    ls -la $O/${TAG}_full_$1.ncu-rep 2>&1 | cut -c20-
  }
  ncu_full conv 'conv3d_pair|conv3d_umma' 6 2 "--import-source on"
  ncu_full wgrad 'wgrad_split_kernel|wgrad_umma_kernel' 3 1 "--import-source on"
  ncu_full stream 'bn_|reparam|cast_f16|pack_dgrad|gate_bwd|wgrad_reduce|wgrad_split_reduce' 33 11 ""
fi

if has net && [ $OK -eq 1 ]; then
  echo "== whole-Net timings (BASELINE.json configs 1-2; not the headline bench)"
  timeout 150 python tools/bench_net.py --steps 3 > $O/${TAG}_bench_net.log 2>&1
  cp $O/bench_net.json $O/${TAG}_bench_net.json 2>/dev/null
  tail -2 $O/${TAG}_bench_net.log | cut -c1-300
fi

if has ref; then
  echo "== reference arm"
  timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>/dev/null
  cut -c1-400 $O/${TAG}_bench_ref.json
fi
echo done
