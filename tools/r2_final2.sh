#!/bin/bash
# Final single-GPU validation and measurements of round 2 (after the diagnostics ring and the
# continuity COLP_OLD request): GPU tests, smoke, bench lines, ncu launch list and one full capture
# of the 8 launches of a moist step.  Usage: gpurun --timeout 1200 -- 'bash tools/r2_final2.sh'
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee gpurun_out/r2f_pytest_gpu.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/r2f_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_cfg4.json 2> gpurun_out/r2f_bench_cfg4.err
for w in "cfg4 --moist" "cfg5" "cfg2" "cfg3" "cfg3 --moist"; do
    name=$(echo $w | tr -d ' -')
    timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --e2e-members 0 > gpurun_out/r2f_bench_$name.json 2> gpurun_out/r2f_bench_$name.err
done
timeout 200 python bench.py --turbulence --steps 5 --warmup 3 --no-cpu-baseline --e2e-members 0 > gpurun_out/r2f_bench_cfg4_turbulence.json 2> gpurun_out/r2f_bench_cfg4_turbulence.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2f_bench_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4g' % d['value'], 'ms %.4g' % d['ms_per_step'],
              'step_roofline', round((d.get('step_roofline') or {}).get('frac', 0), 4),
              'roofline', round((d.get('roofline') or {}).get('frac', 0), 4),
              'e2e %.4g' % d['e2e']['value'], (d.get('cpu_baseline') or {}).get('kind'),
              (d.get('cpu_baseline') or {}).get('value'), d.get('kernel_ms'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
# launch list of the bench command (a number printed under ncu is never a bench value)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    --e2e-members 0 > gpurun_out/r2f_ncu_launches.log 2>&1
# full capture: the 8 launches of one moist Matsuno step
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"k_stage3|k_moist3|k_blocks|k_diag" -s 17 -c 12 -o gpurun_out/r2f_step_moist \
    python tools/kbench.py --moist 1 --steps 1 > gpurun_out/r2f_ncu_full.log 2>&1
ls -la gpurun_out/r2f_* | tail -20
