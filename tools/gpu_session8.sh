#!/bin/bash
# full parity suite, bench, launch list of the bench, ncu --set full of the GEMM shapes and of the batched-query kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/s8_pytest.log 2>&1; echo "pytest exit $?" >> $O/s8_pytest.log; tail -4 $O/s8_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/s8_clocks.csv &
SMI=$!
timeout 900 python bench.py > $O/s8_bench.json 2> $O/s8_bench.err; echo "bench exit $?"
kill $SMI
cat $O/s8_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/s8_bench_ref.json 2>> $O/s8_bench.err; cat $O/s8_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s8_bench_launches.csv python bench.py --steps 2 --warmup 1 > $O/s8_bench_under_ncu.log 2>&1
B=55 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm -c 4 -o $O/s8_gemm_full python tools/gemm_probe.py > $O/s8_ncu_gemm.log 2>&1
REPS=1 BATCH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_batch|k_sort_cand|k_build_shadow|k_gemm' -s 0 -c 12 -o $O/s8_qbatch_full python tools/mem_probe.py > $O/s8_ncu_qbatch.log 2>&1
ls -la $O | tail -12
