#!/bin/bash
# round-1 closing run: full parity suite, smoke, both bench arms, launch list, ncu of the CTA-pair GEMM and the attention kernel
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/rc_pytest.log 2>&1; echo "pytest exit $?" >> $O/rc_pytest.log; tail -3 $O/rc_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 100 > $O/rc_clocks.csv &
SMI=$!
timeout 900 python bench.py > $O/rc_bench.json 2> $O/rc_bench.err; echo "bench exit $?"
kill $SMI
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/rc_bench_ref.json 2>> $O/rc_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/rc_bench_launches.csv python bench.py --steps 2 --warmup 1 > $O/rc_bench_under_ncu.log 2>&1
B=54 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm -c 4 -o $O/rc_gemm_full python tools/gemm_probe.py > $O/rc_ncu_gemm.log 2>&1
B=54 REPS=1 IMPLS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_attention_ws -s 3 -c 1 -o $O/rc_attn_ws_full python tools/attn_time.py > $O/rc_ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'k_preprocess|k_layernorm' -c 3 -o $O/rc_misc_full python tools/attn_probe.py > $O/rc_ncu_misc.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/rc_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','vit_tflops','gpu_launches')}); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['breakdown_ms_per_step']); print('e2e', d['e2e']['value'], d['clocks'])
print({k:(round(v.get('queries_per_s'),1) if isinstance(v,dict) else v) for k,v in d['query'].items() if k!='workload'})
PY
