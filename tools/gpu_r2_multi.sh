#!/bin/bash
# round-2 multi-GPU run: NCCL merge parity, scene-parallel bench (weak) and frame-sharded bench (strong) on N GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; N=${N:-2}; TAG=${TAG:-r2mg}
timeout 600 python -m pytest tests/test_gpu_merge.py -m gpu -q -k "nccl" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-latency --no-wide > $O/${TAG}_weak_${N}gpu.json 2> $O/${TAG}_weak_${N}gpu.err; echo "weak exit $?"
timeout 900 $TR bench.py --gpus $N --sharding frames --steps 2 --warmup 1 > $O/${TAG}_strong_${N}gpu.json 2> $O/${TAG}_strong_${N}gpu.err; echo "strong exit $?"
if [ -n "$STRONG1" ]; then timeout 900 python bench.py --gpus 1 --sharding frames --steps 2 --warmup 1 > $O/${TAG}_strong_1gpu.json 2> $O/${TAG}_strong_1gpu.err; echo "strong1 exit $?"; fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*gpu.json')):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, 'merge', d.get('merge'), 'check', d.get('merge_check'))
PY
tail -3 $O/${TAG}_strong_${N}gpu.err
