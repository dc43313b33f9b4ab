#!/bin/bash
# compute-sanitizer memcheck over the parity tests of the scatter / query / merge / colour kernels and a small ViT forward
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; TAG=${TAG:-san}
for t in ${TESTS:-test_gpu_memory test_gpu_query_batch test_gpu_merge test_gpu_colour test_gpu_cluster}; do
  timeout ${TMO:-500} compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/$t.py -x -q -m gpu -p no:cacheprovider > $O/${TAG}_$t.log 2>&1
  echo "$t exit $? : $(grep -c 'Invalid\|out of bounds\|misaligned' $O/${TAG}_$t.log) findings; $(grep -E 'passed|failed|ERROR SUMMARY' $O/${TAG}_$t.log | tr '\n' ' ')"
done
