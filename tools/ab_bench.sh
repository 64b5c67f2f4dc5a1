#!/usr/bin/env bash
# A/B on ONE box: alternate builds of the library (tools/experiments/ab/*.so, DC_B200_LIB) and environment switches, interleaved.
# usage: tools/ab_bench.sh "<label>=<env assignments>" ...   e.g.  tools/ab_bench.sh "pull=DC_PUSH=0" "B=DC_B200_LIB=tools/experiments/ab/B.so"
cd "$(dirname "$0")/.."
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["launch_ms"])'
for rep in 1 2; do
  for spec in "$@"; do
    label="${spec%%=*}"; envs="${spec#*=}"
    printf "%s rep%s: " "$label" "$rep"
    env $envs timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-conditioning --workload ${WORKLOAD:-C2} 2>&1 | tail -1 | python -c "$J"
  done
done
