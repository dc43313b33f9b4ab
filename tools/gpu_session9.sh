#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
echo "== gemm tests, cta_group::2 on"
BSCNAV_GEMM_CLUSTER=1 timeout 120 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "gemm" 2>&1 | tail -8
echo "== timing cluster=0"; BSCNAV_GEMM_CLUSTER=0 B=55 timeout 200 python tools/perf_probe.py 2>&1 | grep "gemm\|vit forward"
echo "== timing cluster=1 B=54"; BSCNAV_GEMM_CLUSTER=1 B=54 timeout 200 python tools/perf_probe.py 2>&1 | grep "gemm\|vit forward"
echo "== timing cluster=1 B=55"; BSCNAV_GEMM_CLUSTER=1 B=55 timeout 200 python tools/perf_probe.py 2>&1 | grep "gemm\|vit forward"
BSCNAV_GEMM_CLUSTER=1 timeout 300 python -m pytest tests/test_gpu_vit.py tests/test_gpu_query_batch.py -m gpu -x -q 2>&1 | tail -4
