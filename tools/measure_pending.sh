#!/bin/bash
# What round 1 wrote after its last GPU slot and could only check on the host emulation.
# One single-GPU call and one two-GPU call measure / validate all of it:
#
#   gpurun --timeout 420 -- 'bash tools/measure_pending.sh one'
#   gpurun --gpus 2 --timeout 300 -- 'bash tools/measure_pending.sh two'
#
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
case "${1:-one}" in
one)
    # 1. parity of the default paths with the final sources (kernel decomposition with the
    #    coupled terms, turbulence module, streamed e2e)
    python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pending_pytest.log
    # 2. coupled terms beside the fused dry stage kernel:
    #    expected ~18 ms/step against 42.3 ms/step of the kernel decomposition
    python bench.py --turbulence --steps 5 --warmup 3 --e2e-steps 1 \
        > gpurun_out/pending_turb_kernels.json 2> gpurun_out/pending_turb_kernels.err
    DC_COUPLED_IMPL=2 python bench.py --turbulence --steps 5 --warmup 3 --e2e-steps 1 \
        > gpurun_out/pending_turb_fused.json 2> gpurun_out/pending_turb_fused.err
    python - <<'PY'
import json
for f in ('pending_turb_kernels', 'pending_turb_fused'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, '%.2f ms/step' % d['ms_per_step'], 'finite', d['config']['finite'],
              {k: round(v, 2) for k, v in d['kernels_ms_per_step'].items()})
    except Exception as e:
        print(f, 'FAILED', e)
PY
    ;;
two)
    # 3. latitude bands: the solver / output / restart gathers over NCCL and the band-shaped
    #    e2e leg of bench.py (ModelFields.to_device_band / to_host_band)
    python -m pytest tests/test_gpu_bands.py -x -q 2>&1 | tail -4 | tee gpurun_out/pending_bands.log
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 bench.py --gpus 2 > gpurun_out/pending_bench_n2.json \
        2> gpurun_out/pending_bench_n2.err
    python -c "import json; d = json.load(open('gpurun_out/pending_bench_n2.json')); print(d['value'], d['e2e'])"
    grep -h "pre-flight" gpurun_out/pending_bench_n2.err
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29612 -m climate_model_b200.solver --nsteps 20 --output gpurun_out/pending_nc \
        dlat_deg=1.0 dlon_deg=1.0 nz=32 2>&1 | tail -4
    ;;
esac
