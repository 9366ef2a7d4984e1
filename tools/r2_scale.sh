#!/bin/bash
# Round-2 multi-GPU measurements on N GPUs of one box:
#   gpurun --gpus N --timeout 1500 -- 'bash tools/r2_scale.sh N [cfg5]'
# strong scaling of BASELINE configs[3] (0.25 deg x 64 levels) with the in-library exchange, the
# per-stream timeline of one banded step, the bitwise N-GPU == 1-GPU check inside bench.py, and
# (with "cfg5") the weak-scaling workload of configs[4] (0.1 deg x 96 levels, 210 rows per GPU).
N=${1:-2}
mkdir -p gpurun_out
run() {  # name, extra bench args...
    local name=$1; shift
    NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 2 "$@" \
        > gpurun_out/${name}.json 2> gpurun_out/${name}.err
    python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${name}.json'))
    print('${name}', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'],
          'parity', (d.get('parity_vs_1gpu') or {}).get('bitwise_equal'),
          'e2e %.4g' % d['e2e']['value'], 'launches', d['gpu_launches'])
except Exception as e:
    print('${name} FAILED', e)
PY
    grep -h -m1 "Init COMPLETE\|comm 0x.* nranks" gpurun_out/${name}.err | head -2
    tail -2 gpurun_out/${name}.err | cut -c1-300
}
if [ "$N" -gt 1 ]; then
    python -m pytest tests/test_gpu_bands.py -x -q -k "library and $N-" 2>&1 | tail -2
fi
run r2_bench_cfg4_n$N --timeline gpurun_out/r2_timeline_cfg4_n$N.json
if [ "${2:-}" = "cfg5" ]; then
    run r2_bench_cfg5_n$N --workload cfg5
fi
