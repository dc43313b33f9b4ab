#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_cluster.py -m gpu -x -q 2>&1 | tail -6
echo "== gemm tests, cluster multicast on"
BSCNAV_GEMM_CLUSTER=1 timeout 300 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "gemm" 2>&1 | tail -6
echo "== timing cluster=0"; BSCNAV_GEMM_CLUSTER=0 B=55 timeout 300 python tools/perf_probe.py 2>&1 | grep "gemm\|vit forward"
echo "== timing cluster=1"; BSCNAV_GEMM_CLUSTER=1 B=55 timeout 300 python tools/perf_probe.py 2>&1 | grep "gemm\|vit forward"
echo "== timing cluster=1 B=54"; BSCNAV_GEMM_CLUSTER=1 B=54 timeout 300 python tools/perf_probe.py 2>&1 | grep "gemm\|vit forward"
BSCNAV_GEMM_CLUSTER=1 timeout 600 python -m pytest tests/test_gpu_vit.py tests/test_gpu_query_batch.py -m gpu -x -q 2>&1 | tail -4
