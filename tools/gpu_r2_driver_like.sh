#!/bin/bash
# what the driver runs for the scaling table: bench.py with DEFAULT flags under torchrun on N GPUs, own arm then reference arm;
# then the frame-sharded (strong) mode with its e2e leg
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; N=${N:-2}; TAG=${TAG:-r2dl}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_default_${N}gpu.json 2> $O/${TAG}_default_${N}gpu.err; echo "default exit $?"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/${TAG}_ref_${N}gpu.json 2> $O/${TAG}_ref_${N}gpu.err; echo "reference exit $?"
timeout 900 $TR bench.py --gpus $N --sharding frames --steps 2 --warmup 1 > $O/${TAG}_strong_${N}gpu.json 2> $O/${TAG}_strong_${N}gpu.err; echo "strong exit $?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*_${N}gpu.json')):
    try:
        lines=[l for l in open(f).read().splitlines() if l.startswith('{')]
        d=json.loads(lines[-1]); print(f, 'json lines', len(lines))
    except Exception as e:
        print(f, 'unreadable', e); continue
    print('  ', {k:d.get(k) for k in ('impl','value','ms_per_step','n_gpus','scaling')}, 'e2e', (d.get('e2e') or {}).get('value'), 'merge', d.get('merge'))
PY
for f in default strong; do tail -n 2 $O/${TAG}_${f}_${N}gpu.err; done; exit 0
