#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python -m pytest tests/test_gpu_vit.py -m gpu -x -q -k "attention" 2>&1 | tail -3
B=54 IMPLS=0,2 timeout 200 python tools/attn_time.py
timeout 200 python tools/attn_timeline.py | head -8
