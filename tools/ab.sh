#!/bin/bash
# Same-box A/B of library builds: tools/ab.sh ab/lib_old.so ab/lib_new.so [rounds]
R=${3:-3}
for i in $(seq $R); do
  for L in "$1" "$2"; do
    echo -n "$L "; VADB200_LIB=$PWD/$L python bench.py --hours-per-gpu 400 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*2.5,2))"
  done
done
