#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "gemm or full_vit or tiny" 2>&1 | tail -3
B=54 timeout 200 python tools/perf_probe.py 2>&1 | grep "act=0"
timeout 300 python bench.py --model dinov2 --steps 16 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('dinov2', round(d['value'],1), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1), d['roofline']['breakdown_ms_per_step'])"
