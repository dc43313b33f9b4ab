#!/bin/bash
# closing run of a session: full parity suite, smoke, default bench (no profiler)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/fc_pytest.log 2>&1; echo "pytest exit $?" >> $O/fc_pytest.log; tail -3 $O/fc_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $O/fc_bench.json 2> $O/fc_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/fc_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline'].get('breakdown_ms_per_step')); print('e2e', d['e2e']['value'], d['clocks'])
print({k:(round(v.get('queries_per_s'),1) if isinstance(v,dict) and 'queries_per_s' in v else '') for k,v in d['query'].items() if k!='workload'})
PY
