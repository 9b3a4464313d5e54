export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "3d or awkward or generations" > gpurun_out/t.log 2>&1; tail -n 3 gpurun_out/t.log
run() { # label, env..., args
  label=$1; shift
  env "$@" timeout 200 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 $ARGS > gpurun_out/x.json 2> gpurun_out/x.err || { echo "$label FAILED"; tail -n 3 gpurun_out/x.err; return; }
  echo "$label: $(python tools/show_bench.py gpurun_out/x.json | cut -c1-60)"
}
ARGS="--dim 3 --n 256"
run "256 zpb32" A=1
run "256 zpb16" SB2_RD_ZPB=16
ARGS="--dim 3 --n 384"
run "384 zpb32" A=1
run "384 zpb16" SB2_RD_ZPB=16
run "384 zpb64" SB2_RD_ZPB=64
ARGS="--dim 3"
run "512 zpb32" A=1
