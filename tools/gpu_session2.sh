#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "attention" > $O/s2_pytest_attn.log 2>&1; echo "exit $?" >> $O/s2_pytest_attn.log
tail -15 $O/s2_pytest_attn.log
B=55 IMPLS=0,1,2 timeout 200 python tools/attn_time.py > $O/s2_attn_time.log 2>&1
B=55 T=261 IMPLS=0,2 timeout 200 python tools/attn_time.py >> $O/s2_attn_time.log 2>&1
cat $O/s2_attn_time.log
timeout 900 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "full_vit" > $O/s2_pytest_vit.log 2>&1; echo "exit $?" >> $O/s2_pytest_vit.log
tail -5 $O/s2_pytest_vit.log
BSCNAV_ATTENTION=2 timeout 600 python bench.py > $O/s2_bench_ws.json 2> $O/s2_bench.err; cat $O/s2_bench_ws.json
B=55 REPS=1 IMPLS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_attention_ws -s 3 -c 1 -o $O/s2_attn_ws_full python tools/attn_time.py > $O/s2_ncu_attn.log 2>&1
