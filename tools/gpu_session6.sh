#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_query_batch.py -m gpu -x -q > $O/s6_pytest.log 2>&1; echo "exit $?" >> $O/s6_pytest.log
tail -25 $O/s6_pytest.log
REPS=5 timeout 300 python tools/mem_probe.py > $O/s6_mem_probe.log 2>&1; cat $O/s6_mem_probe.log
