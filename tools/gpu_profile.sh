#!/bin/bash
# Run on the GPU box (under gpurun): parity tests, the default bench line, the
# ncu launch list of the same command and one full capture of the fused kernel.
#   tools/gpu_profile.sh <tag> [bench args...]
# Outputs land in gpurun_out/<tag>_*; summarise them here with tools/ncu_summary.py.
set -u
tag=${1:-r01}; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${tag}_clocks.csv &
smi=$!
python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
kill $smi
cat gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eval-shape --no-train-step "$@" > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gqe_fused -s 4 -c 1 -f -o gpurun_out/${tag}_fused_full \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eval-shape --no-train-step "$@" > gpurun_out/${tag}_full.log 2>&1
ls -la gpurun_out | tail -12
