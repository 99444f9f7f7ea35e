#!/bin/bash
# Round evidence in one gpurun call: final bench lines, ncu launch list, ncu --set full captures of the three
# fused-kernel modes and the streaming kernel.  Every ncu command runs only after the identical plain command exited 0.
set -u
O=gpurun_out
T=${1:-r2}
python bench.py > $O/${T}_final_bench.json 2> $O/${T}_final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_reference_arm.json 2> $O/${T}_reference_arm.err
L="python bench.py --hours-per-gpu 100 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-extra-configs"
$L > $O/${T}_launch_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv $L > $O/${T}_launch_ncu.log 2>&1
K="python bench.py --hours-per-gpu 50 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs"
$K > $O/${T}_k22_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 3 -c 1 -o $O/${T}_prof_k22 $K > $O/${T}_k22_ncu.log 2>&1
for C in cfg2 cfg5; do
  python tools/bench_configs.py $C > $O/${T}_${C}_plain.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 4 -c 1 -o $O/${T}_prof_$C python tools/bench_configs.py $C > $O/${T}_${C}_ncu.log 2>&1
done
S="python tools/bench_configs.py stream --stream-ticks 300"
$S > $O/${T}_stream_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:stream_feed -s 50 -c 1 -o $O/${T}_prof_stream $S > $O/${T}_stream_ncu.log 2>&1
ls -la $O | tail -30
