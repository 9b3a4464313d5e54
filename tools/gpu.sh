#!/bin/bash
# usage: tools/gpu.sh <timeout-seconds> '<command run on the GPU box>'  — rebuilds the library first, so that the
# snapshot sent to the box never carries a stale .so
set -e
cd "$(dirname "$0")/.."
bash build.sh > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; exit 1; }
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
