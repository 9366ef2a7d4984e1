#!/bin/bash
# final single-GPU bench lines of round 2 (profiles/r2_bench_*.json)
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
for w in "cfg4 --moist" "cfg5" "cfg2" "cfg3"; do
    name=$(echo $w | tr -d ' -')
    python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --e2e-members 0 > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4g' % d['value'], 'ms %.4g' % d['ms_per_step'],
              'step_roofline', round((d.get('step_roofline') or {}).get('frac', 0), 4),
              'e2e %.4g' % d['e2e']['value'], (d.get('cpu_baseline') or {}).get('kind'),
              (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
