#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_gpu_vit.py tests/test_gpu_facade.py tests/test_gpu_query_batch.py -m gpu -x -q 2>&1 | tail -3
for pdl in 1 0 1 0; do
BSCNAV_PDL=$pdl timeout 300 python bench.py --steps 32 --warmup 4 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('PDL=$pdl', round(d['value'],1), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1))"
done
