#!/bin/bash
# usage (on the GPU box, via tools/gpu.sh): tools/profile.sh <tag>  — bench line, ncu launch list of the same command,
# ncu full capture of the four stage kernels of one step
tag=${1:-x}
export PYTHONUNBUFFERED=1
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { tail -5 gpurun_out/bench_$tag.err; exit 1; }
python tools/show_bench.py gpurun_out/bench_$tag.json
CMD="python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/plain_$tag.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncul_$tag.log 2>&1
timeout 200 $CMD > gpurun_out/plain_$tag.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k "regex:stage_kernel|rd_jit" -s 12 -c 4 -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_$tag.log 2>&1
tail -n 1 gpurun_out/ncu_$tag.log | cut -c1-120
