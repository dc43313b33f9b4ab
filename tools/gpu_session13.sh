#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_gpu_colour.py tests/test_gpu_facade.py tests/test_gpu_memory.py -m gpu -x -q 2>&1 | tail -15
