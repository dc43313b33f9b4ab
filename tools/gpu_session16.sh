#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "attention or full_vit" 2>&1 | tail -4
B=54 IMPLS=0,2 timeout 200 python tools/attn_time.py
B=54 T=261 IMPLS=0,2 timeout 200 python tools/attn_time.py
B=54 REPS=1 IMPLS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_attention_ws -s 3 -c 1 -o $O/s16_attn_ws_full python tools/attn_time.py > $O/s16_ncu_attn.log 2>&1
