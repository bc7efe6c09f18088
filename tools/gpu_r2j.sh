#!/bin/bash
# r2j (1 GPU): K1b reading K4's partials: parity tests, headline line (+ configs 2 / 3), net_train with the path off.
TAG=${1:-r2j}
O=gpurun_out; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1 REPMODE_NO_BUILD=1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest.log | tail -12 | cut -c1-300
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'sustained', d['sustained']['ms_per_step'], 'launches', d['gpu_launches'])
print({k: (v.get('ms_per_step') if isinstance(v, dict) else v) for k, v in (d.get('other_configs') or {}).items()})
PY
tail -2 $O/${TAG}_bench.err | cut -c1-300
REPMODE_K1B_FROM_PARTIALS=0 timeout 200 python bench.py --config net_train --steps 10 --warmup 3 > $O/${TAG}_net_train_off.json 2> $O/${TAG}_net_train_off.err
echo "net_train with d_weff: $(grep -o '"ms_per_step": [0-9.]*' $O/${TAG}_net_train_off.json | head -1)"
echo done
