#!/bin/bash
# round-2 check run: GPU parity suite, smoke, both bench arms.  TAG names the outputs (gpurun_out/$TAG_*)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; TAG=${TAG:-r2}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -15 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py ${BENCH_ARGS} > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"; tail -5 $O/${TAG}_bench.err
if [ -n "$REF" ]; then timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; echo "ref exit $?"; fi
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','vit_tflops','gpu_launches')}); r=d['roofline']; print(r['achieved'], r['frac'], r.get('achieved_min_max'), r['breakdown_ms_per_step'], r['whole_step'])
print('e2e', d['e2e'], d['clocks']); print('scatter', d.get('scatter')); print('latency', d.get('latency_b1')); print('cpu', d.get('cpu_baseline'))
print({k:(round(v.get('queries_per_s'),1) if isinstance(v,dict) and 'queries_per_s' in v else None) for k,v in d['query'].items()})
PY
if [ -n "$REF" ]; then cat $O/${TAG}_bench_ref.json; fi
