#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/s10_pytest.log 2>&1; echo "pytest exit $?" >> $O/s10_pytest.log; tail -4 $O/s10_pytest.log
timeout 900 python bench.py > $O/s10_bench.json 2> $O/s10_bench.err; echo "bench exit $?"; tail -3 $O/s10_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s10_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','vit_tflops','gpu_launches')}); print(d['roofline']['achieved'], d['roofline']['breakdown_ms_per_step']); print(d['e2e']['value'], d['clocks'])
PY
timeout 600 python bench.py --batch 55 --steps 8 > $O/s10_bench_b55.json 2>> $O/s10_bench.err; python -c "
import json;d=json.load(open('gpurun_out/s10_bench_b55.json'));print('B=55', d['value'], d['ms_per_step'])"
BSCNAV_GEMM_CLUSTER=0 timeout 600 python bench.py --batch 55 --steps 8 > $O/s10_bench_nocl.json 2>> $O/s10_bench.err; python -c "
import json;d=json.load(open('gpurun_out/s10_bench_nocl.json'));print('no-cluster B=55', d['value'], d['ms_per_step'])"
