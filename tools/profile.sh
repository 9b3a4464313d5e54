#!/bin/bash
# usage (on the GPU box, via gpurun): tools_profile.sh <tag>  — parity tests, bench line, ncu full capture of the stage kernel
tag=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
print('ms/step',d['ms_per_step'],'value %.3e'%d['value'],'frac',d['roofline']['frac'],'e2e %.3e'%d['e2e']['value'],'clk',d['clocks'])
" || tail -5 gpurun_out/bench_$tag.err
python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 12 -c 4 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
tail -1 gpurun_out/ncu_$tag.log
