#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { # name env...
  name=$1; shift
  for rep in 1 2; do
    env "$@" timeout 600 python bench.py --steps 48 --warmup 5 $EXTRA > $O/s11_tmp.json 2>> $O/s11.err
    python -c "
import json;d=json.load(open('gpurun_out/s11_tmp.json'));print('$name rep$rep', round(d['value'],1), 'fps', round(d['ms_per_step'],3), 'ms/step  e2e', round(d['e2e']['value'],1), ' gemm TF', round(d['roofline']['achieved'],1))"
  done
}
EXTRA="--batch 54" run "cluster1 overlap1 B54" BSCNAV_GEMM_CLUSTER=1 BSCNAV_OVERLAP=1
EXTRA="--batch 54" run "cluster1 overlap0 B54" BSCNAV_GEMM_CLUSTER=1 BSCNAV_OVERLAP=0
EXTRA="--batch 55" run "cluster0 overlap1 B55" BSCNAV_GEMM_CLUSTER=0 BSCNAV_OVERLAP=1
EXTRA="--batch 55" run "cluster0 overlap0 B55" BSCNAV_GEMM_CLUSTER=0 BSCNAV_OVERLAP=0
EXTRA="--batch 108" run "cluster1 overlap1 B108" BSCNAV_GEMM_CLUSTER=1 BSCNAV_OVERLAP=1
