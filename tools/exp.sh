export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t.log 2>&1; tail -n 3 gpurun_out/t.log
run() { # label, env..., args
  label=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 $ARGS > gpurun_out/x_$label.json 2> gpurun_out/x.err || { echo "$label FAILED"; tail -n 3 gpurun_out/x.err; return; }
  echo "$label: $(python tools/show_bench.py gpurun_out/x_$label.json | cut -c1-60)"
}
ARGS="--dim 3 --n 384"
run "3d384" A=1
ARGS="--dim 3"
run "3d512" A=1
ARGS=""
run "2d" A=1
