#!/usr/bin/env bash
# The first gpurun call of the next round: everything written after the GPU budget of round 2 was spent
#  has never executed on hardware.  Run from the repo root:
#
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/first_gpu_call_next_round.sh'
#
# 1. the late-sorting GPU tests alone (composition of verified kernels, lazy file sources, the fire-weather kernel),
# 2. the fire-weather kernel timed on a lat band with the oracle leg beside it,
# 3. its launch list and one full ncu capture (FP64 pipe share, registers, local memory of the RINGS variant).
set -x
mkdir -p gpurun_out
python -m pytest tests/test_zzx_gpu_late.py tests/test_zzy_gpu_io.py tests/test_zzz_gpu_fire.py -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_late.log
python bench_extra.py --lat 90 --steps 3 --only fwi --cpu > gpurun_out/extra_fwi.jsonl 2> gpurun_out/extra_fwi.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_fwi.csv \
    python bench_extra.py --lat 32 --steps 1 --warmup 0 --only fwi > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fwi_kernel -c 1 -o gpurun_out/fwi_full \
    python bench_extra.py --lat 32 --steps 1 --warmup 0 --only fwi > /dev/null 2>&1
python tools/ncu_summarise.py gpurun_out/fwi_full.ncu-rep > gpurun_out/ncu_fwi.md 2>&1 || true
tail -5 gpurun_out/pytest_late.log
cat gpurun_out/extra_fwi.jsonl
