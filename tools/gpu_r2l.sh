#!/bin/bash
# r2l (1 GPU): K1 with the gate softmax overlapped, eval affine cache; ncu of K1 / K1b on the 512 -> 512 layer
TAG=${1:-r2l}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_net.py -m gpu -q -p no:cacheprovider --timeout 400 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -6 | cut -c1-300
timeout 200 python tools/bench_k1.py > $O/${TAG}_bench_k1.txt 2>&1; cat $O/${TAG}_bench_k1.txt | grep -v Warn
timeout 300 python bench.py --config net_fwd --steps 20 --warmup 5 > $O/${TAG}_net_fwd.json 2> $O/${TAG}_net_fwd.err
echo "net_fwd: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_fwd.json | head -1)"; tail -2 $O/${TAG}_net_fwd.err | cut -c1-200
K1_ONLY=512 timeout 200 ncu --set full --clock-control none --import-source on -k regex:'reparam_bwd_kernel|reparam_fwd_rows_kernel|pack_dgrad_kernel' -s 9 -c 3 \
  -o $O/${TAG}_full_k1 -f python tools/bench_k1.py > $O/${TAG}_ncu_k1.log 2>&1
[ -s $O/${TAG}_full_k1.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_k1.ncu-rep --page raw --csv > $O/${TAG}_full_k1_raw.csv 2>/dev/null
[ -s $O/${TAG}_full_k1.ncu-rep ] && timeout 30 ncu -i $O/${TAG}_full_k1.ncu-rep --page source --csv > $O/${TAG}_full_k1_src.csv 2>/dev/null
ls -la $O/${TAG}_full_k1* | cut -c20-
echo done
