#!/usr/bin/env bash
# One GPU visit: parity tests, smoke, bench (both arms), ncu launch list + full capture of the persistent clip kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
if [ "${TRY_INVERT:-0}" = "1" ]; then echo "=== pytest small with DC_MASK_INVERT=1"; DC_MASK_INVERT=1 timeout 300 python -m pytest tests -x -q -m gpu -k "forward_small" 2>&1 | tail -5; fi
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_ours.json
echo "=== bench C3"; timeout 900 python bench.py --steps 3 --warmup 3 --workload C3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_c3.json
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-conditioning > gpurun_out/ncu_launch_run.log 2>&1
tail -3 gpurun_out/ncu_launch_run.log
echo "=== ncu full (layer kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(clip_kernel|layer_kernel)" -s 3 -c 1 -f -o gpurun_out/clip_full \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-conditioning > gpurun_out/ncu_full_run.log 2>&1
tail -3 gpurun_out/ncu_full_run.log
ls -la gpurun_out | head -30
fi
