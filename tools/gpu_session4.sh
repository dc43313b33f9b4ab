#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "attention" > $O/s4_pytest_attn.log 2>&1; echo "exit $?" >> $O/s4_pytest_attn.log
tail -4 $O/s4_pytest_attn.log
BSCNAV_ATTN_PINGPONG=1 timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "attention" > $O/s4_pytest_attn_pp.log 2>&1; echo "exit $?" >> $O/s4_pytest_attn_pp.log
tail -4 $O/s4_pytest_attn_pp.log
B=55 IMPLS=0,2 timeout 200 python tools/attn_time.py > $O/s4_attn_time.log 2>&1
BSCNAV_ATTN_PINGPONG=1 B=55 IMPLS=2 timeout 200 python tools/attn_time.py >> $O/s4_attn_time.log 2>&1
B=55 T=261 IMPLS=0,2 timeout 200 python tools/attn_time.py >> $O/s4_attn_time.log 2>&1
cat $O/s4_attn_time.log
B=55 REPS=1 IMPLS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_attention_ws -s 3 -c 1 -o $O/s4_attn_ws_full python tools/attn_time.py > $O/s4_ncu_attn.log 2>&1
