#!/bin/bash
# bench.py (no CPU baseline) under a list of environment settings: "NAME|VAR=VAL VAR=VAL"
mkdir -p gpurun_out
: > gpurun_out/env_matrix.txt
for spec in "$@"; do
  name="${spec%%|*}"; envs="${spec#*|}"
  out=$(env $envs timeout 100 python bench.py --no-cpu-baseline 2>gpurun_out/env_${name}.err | grep '^{')
  echo "$name|$envs|$(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read() or '{}'); print(round(d.get('value',0),1), round(d.get('e2e',{}).get('value',0),1), d.get('ms_per_step'))")" | tee -a gpurun_out/env_matrix.txt
done
