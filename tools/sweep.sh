#!/bin/bash
# usage: tools/sweep.sh TAG "ENV=VAL ..."   -> one bench run, prints seq-evals/s
TAG=$1; shift
env "$@" python bench.py --nseq 10000 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/sweep_$TAG.log 2>&1
python - gpurun_out/sweep_$TAG.log $TAG <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); print(sys.argv[2], round(d["seq_evals_per_s"]), round(d["ms_per_step"], 1))
PY
