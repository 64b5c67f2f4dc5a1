#!/usr/bin/env bash
# One GPU visit: parity tests, smoke, bench (both arms, every workload that fits one GPU), ncu launch list + full capture of the
# persistent clip kernel.  SKIP_TESTS=1 / SKIP_NCU=1 / SKIP_EXTRA=1 shorten it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
fi
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_ours.json
echo "=== bench C3"; timeout 900 python bench.py --steps 3 --warmup 3 --workload C3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_c3.json
if [ "${SKIP_EXTRA:-0}" != "1" ]; then
echo "=== bench C4 (512 clips on one GPU)"; timeout 900 python bench.py --steps 3 --warmup 3 --workload C4 --no-cpu-baseline --no-conditioning 2>&1 | tail -2 | tee gpurun_out/bench_c4.json
echo "=== bench C5 (1000-step DDPM, 32 x 1800 frames)"; timeout 900 python bench.py --steps 1 --warmup 3 --workload C5 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_c5.json
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-conditioning > gpurun_out/ncu_launch_run.log 2>&1
tail -3 gpurun_out/ncu_launch_run.log
echo "=== ncu full (clip kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(clip_kernel|layer_kernel)" -s 3 -c 1 -f -o gpurun_out/clip_full \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-conditioning > gpurun_out/ncu_full_run.log 2>&1
tail -3 gpurun_out/ncu_full_run.log
ls -la gpurun_out | head -30
fi
