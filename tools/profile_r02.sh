#!/bin/bash
# usage (on the GPU box, via tools/gpu.sh): tools/profile_r02.sh <tag>
# the default bench line (all secondary workloads); then, each only after its command ran clean without ncu: the ncu launch
# list of the bench command, full captures of the stage kernels at 8192^2 and 512^3 and of the CIP-CSLR passes at 4096^2
tag=${1:-r02}
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 python bench.py > $O/bench_$tag.json 2> $O/bench_$tag.err || { tail -5 $O/bench_$tag.err; exit 1; }
python tools/show_bench.py $O/bench_$tag.json
CMD="python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-secondary --repeats 1"
timeout 300 $CMD > $O/plain_$tag.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_$tag.csv $CMD > $O/ncul_$tag.log 2>&1
timeout 300 $CMD > $O/plain_$tag.log 2>&1 && timeout 500 ncu --set full --clock-control none --import-source on -k "regex:rd_jit" -s 12 -c 4 -o $O/prof_${tag}_2d $CMD > $O/ncu_${tag}_2d.log 2>&1
CMD3="python bench.py --dim 3 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-secondary --repeats 1"
timeout 300 $CMD3 > $O/plain3_$tag.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches3d_$tag.csv $CMD3 > $O/ncul3_$tag.log 2>&1
timeout 300 $CMD3 > $O/plain3_$tag.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k "regex:rd_jit" -s 12 -c 4 -o $O/prof_${tag}_3d $CMD3 > $O/ncu_${tag}_3d.log 2>&1
CMDA="python tools/adv_bench.py 4096 3"
timeout 200 $CMDA > $O/adv_$tag.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k "regex:cip2d" -s 4 -c 4 -o $O/prof_${tag}_cip $CMDA > $O/ncu_${tag}_cip.log 2>&1
python tools/adv_bench.py 4096 10; python tools/adv_bench.py 8192 5
tail -n 1 $O/ncu_${tag}_cip.log | cut -c1-120
