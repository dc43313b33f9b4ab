#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
REPS=2 BATCH=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_gemm|k_batch|k_topk|k_row_ok|k_prep|k_build|k_mark|k_query' --csv --log-file $O/s7_query_launches.csv python tools/mem_probe.py > $O/s7_probe.log 2>&1
tail -3 $O/s7_probe.log
