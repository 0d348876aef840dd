#!/bin/bash
# usage: tools/sweep_env.sh TAG VAR v1 v2 ...   -- one short bench.py run per value of an environment switch, key figures to gpurun_out/TAG_*.log
tag=$1; var=$2; shift 2
for v in "$@"; do
  env $var=$v python bench.py --no-cpu-baseline --steps 30 --warmup 3 > gpurun_out/${tag}_${var}_$v.log 2>&1
  echo "$var=$v rc=$? $(grep -o '"share_of_step": {[^}]*}' gpurun_out/${tag}_${var}_$v.log) $(grep -o '"frac": [0-9.]*' gpurun_out/${tag}_${var}_$v.log | head -1) e2e $(grep -o '"e2e": {[^}]*}' gpurun_out/${tag}_${var}_$v.log | grep -o '"ms_per_step": [0-9.]*')"
done
