#!/bin/bash
# r2f (1 GPU): the round's evidence set -- parity tests, the driver's bench line (+ reference arm), ncu launch list of the
# headline step, capped `ncu --set full` captures (conv pair, wgrad deep, K1, streaming kernels), net_fwd / net_train lines.
# Every step under its own short timeout; ncu always with -k and -c.
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?"; cut -c1-600 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
echo "ref exit $?"; cut -c1-400 $O/${TAG}_bench_ref.json
timeout 120 python tools/step_breakdown.py > $O/${TAG}_breakdown.log 2>&1
cp $O/step_breakdown.json $O/${TAG}_step_breakdown.json 2>/dev/null; tail -2 $O/${TAG}_breakdown.log | cut -c1-300

echo "== ncu launch list"
REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none \
  -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_list.log 2>&1
wc -l $O/${TAG}_launches.csv
ncu_full() {   # name, kernel regex, skip, count, extra flags
  REPMODE_OVERLAP=0 REPMODE_BENCH_FAST=2 REPMODE_BENCH_GRAPH=0 timeout 150 ncu --set full --clock-control none $5 \
    -k regex:"$2" -s $3 -c $4 -o $O/${TAG}_full_$1 -f python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_full_$1.log 2>&1
  [ -s $O/${TAG}_full_$1.ncu-rep ] && timeout 60 ncu -i $O/${TAG}_full_$1.ncu-rep --page raw --csv > $O/${TAG}_full_$1_raw.csv 2>/dev/null
  ls -la $O/${TAG}_full_$1.ncu-rep 2>&1 | cut -c20-
}
ncu_full conv 'conv3d_pair' 6 2 "--import-source on"
ncu_full wgrad 'wgrad_deep' 4 2 "--import-source on"
ncu_full stream 'bn_|reparam|cast_f16|pack_dgrad|gate_bwd|amax' 30 12 ""
[ -s $O/${TAG}_full_wgrad.ncu-rep ] && timeout 60 ncu -i $O/${TAG}_full_wgrad.ncu-rep --page source --csv > $O/${TAG}_full_wgrad_src.csv 2>/dev/null
[ -s $O/${TAG}_full_conv.ncu-rep ] && timeout 60 ncu -i $O/${TAG}_full_conv.ncu-rep --page source --csv > $O/${TAG}_full_conv_src.csv 2>/dev/null
rm -f $O/${TAG}_full_stream.ncu-rep

timeout 200 python tools/bench_k1.py > $O/${TAG}_bench_k1.txt 2>&1; grep -v Warn $O/${TAG}_bench_k1.txt | tail -12
timeout 300 python bench.py --config net_fwd --steps 20 --warmup 5 > $O/${TAG}_net_fwd.json 2> $O/${TAG}_net_fwd.err
echo "net_fwd: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_fwd.json | head -1)"
timeout 400 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train.json 2> $O/${TAG}_net_train.err
echo "net_train: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_train.json | head -1) launches $(grep -o '"gpu_launches": [0-9]*' $O/${TAG}_net_train.json)"; tail -3 $O/${TAG}_net_train.err
timeout 300 python tools/profile_net.py --train --batch 4 > $O/${TAG}_profile_net_train4.txt 2>&1
timeout 200 python tools/profile_net.py > $O/${TAG}_profile_net_eval.txt 2>&1
echo done
