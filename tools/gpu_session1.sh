#!/bin/bash
# GPU session: parity suite, attention timing, bench, ncu captures of the HBM kernels and the attention kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/s1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/s1_pytest.log 2>&1; echo "pytest exit $?" >> $O/s1_pytest.log
B=55 timeout 200 python tools/attn_time.py > $O/s1_attn_time.log 2>&1
timeout 300 python tools/mem_probe.py > $O/s1_mem_probe.log 2>&1
VOXEL=1 timeout 300 python tools/mem_probe.py >> $O/s1_mem_probe.log 2>&1
timeout 600 python bench.py > $O/s1_bench.json 2> $O/s1_bench.err
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_index|k_select|k_commit|k_count|k_scan|k_assign|k_query|k_topk' -c 14 -o $O/s1_mem_full python tools/mem_probe.py > $O/s1_ncu_mem.log 2>&1
B=55 REPS=1 IMPLS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_attention_tc -s 3 -c 1 -o $O/s1_attn_tc_full python tools/attn_time.py > $O/s1_ncu_attn.log 2>&1
B=55 REPS=1 IMPLS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_attention -s 3 -c 1 -o $O/s1_attn_mma_full python tools/attn_time.py >> $O/s1_ncu_attn.log 2>&1
tail -3 $O/s1_pytest.log; cat $O/s1_attn_time.log $O/s1_mem_probe.log; cat $O/s1_bench.json
