export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t.log 2>&1; tail -n 3 gpurun_out/t.log
