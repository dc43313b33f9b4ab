#!/bin/bash
# round-2 ncu captures (--set full, one GPU): attention (second generation), the four GEMM shapes, the query path, the scatter path
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
B=54 REPS=1 IMPLS=3 timeout 300 $NCU -k regex:k_attention_ws2 -s 3 -c 1 -f -o $O/r02_attention_ws2_full python tools/attn_time.py > $O/r02_ncu_attn.log 2>&1; echo "attn $?"
B=54 ITERS=1 timeout 600 $NCU -k regex:k_gemm -c 4 -f -o $O/r02_gemm_full python tools/gemm_probe.py > $O/r02_ncu_gemm.log 2>&1; echo "gemm $?"
timeout 600 $NCU -k regex:'k_query_slots|k_keys|k_sort|k_topk_decode|k_gemm|k_blk|k_batch_rescore|k_build_shadow' -c 14 -f -o $O/r02_query_full python tools/query_probe.py > $O/r02_ncu_query.log 2>&1; echo "query $?"
REPS=1 timeout 600 $NCU -k regex:'k_index|k_count|k_scan|k_assign|k_select|k_commit' -s 30 -c 6 -f -o $O/r02_mem_full python tools/mem_probe.py > $O/r02_ncu_mem.log 2>&1; echo "mem $?"
timeout 120 tools/micro/mma_chain > $O/r02_mma_issue_rate.txt 2>&1; tail -3 $O/r02_mma_issue_rate.txt
ls -la $O/r02_*.ncu-rep
