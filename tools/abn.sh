#!/bin/bash
# Same-box comparison of several library builds: tools/abn.sh ROUNDS lib1.so lib2.so ...   (ms per 1000 h)
R=$1; shift
for i in $(seq $R); do
  for L in "$@"; do
    echo -n "$L "; VADB200_LIB=$PWD/$L python bench.py --hours-per-gpu 400 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*2.5,2))"
  done
done
