#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "preprocess or full_vit or tiny" 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > $O/s14_bench.json 2> $O/s14_bench.err; echo "bench exit $?"; tail -3 $O/s14_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s14_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','vit_tflops','gpu_launches')}); print(d['roofline']['achieved'], d['roofline']['breakdown_ms_per_step']); print('e2e', d['e2e']['value'], d['clocks'])
print({k:(v.get('queries_per_s') if isinstance(v,dict) else v) for k,v in d['query'].items()})
PY
