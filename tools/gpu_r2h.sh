#!/bin/bash
# r2h (1 GPU): K4 in two phases with K3 forked in between; eval-cache invalidation hooks: parity tests, smoke, headline line,
# net_train, and the D-sharded formulation on one GPU.
TAG=${1:-r2h}
O=gpurun_out; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1 REPMODE_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/${TAG}_smoke.log | cut -c1-300
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'sustained', d['sustained']['ms_per_step'], 'launches', d['gpu_launches'])
print('roofline frac', d['roofline']['frac'], 'wgrad', d['roofline']['slowest_kernel']['frac'], 'step', d['roofline']['whole_step']['frac'])
print({k: (v.get('ms_per_step') if isinstance(v, dict) else v) for k, v in (d.get('other_configs') or {}).items()})
PY
tail -3 $O/${TAG}_bench.err | cut -c1-300
REPMODE_WGRAD_PHASES=0 REPMODE_BENCH_CONFIGS=0 timeout 300 python bench.py > $O/${TAG}_bench_onephase.json 2> $O/${TAG}_bench_onephase.err
echo "one-phase: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_bench_onephase.json | head -1)"
echo done
