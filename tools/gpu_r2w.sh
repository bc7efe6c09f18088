#!/bin/bash
# r2w (1 GPU): launch-structure changes (finalize folded into the apply kernel, dy scale inside the backward apply kernel,
# zero buffers / step counter on the side stream, cached sample indices, block-form gate backward): parity tests, the
# headline line (with configs 2 / 3 from child runs), step breakdown, and a 1-GPU pass over config 5's per-rank slab shape.
TAG=${1:-r2w}
O=gpurun_out
mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
export REPMODE_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'sustained', d['sustained']['ms_per_step'], 'launches', d['gpu_launches'])
print('roofline frac', d['roofline']['frac'], 'wgrad', d['roofline']['slowest_kernel']['frac'], 'step', d['roofline']['whole_step']['frac'])
print(json.dumps(d.get('other_configs'))[:900])
PY
tail -3 $O/${TAG}_bench.err | cut -c1-300
timeout 120 python tools/step_breakdown.py > $O/${TAG}_breakdown.log 2>&1
cp $O/step_breakdown.json $O/${TAG}_step_breakdown.json 2>/dev/null; tail -2 $O/${TAG}_breakdown.log | cut -c1-300
timeout 200 python tools/bench_k1.py > $O/${TAG}_bench_k1.txt 2>&1; grep -v Warn $O/${TAG}_bench_k1.txt | tail -7
REPMODE_BENCH_CFG_DIMS=16,512,512 timeout 300 python bench.py --config cfg5 --steps 3 --warmup 2 > $O/${TAG}_cfg5_slab_1gpu.json 2> $O/${TAG}_cfg5_slab_1gpu.err
echo "cfg5 slab shape on 1 GPU exit $?"; grep -o '"ms_per_step": [0-9.]*\|"mem_gb": [0-9.]*' $O/${TAG}_cfg5_slab_1gpu.json | head -3 | tr '\n' ' '; tail -2 $O/${TAG}_cfg5_slab_1gpu.err | cut -c1-300
echo done
