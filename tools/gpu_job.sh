#!/bin/bash
# Run a command on the GPU box from a FROZEN copy of the tree (.frozen/cur, taken now), so that the working tree can keep
# changing while the job waits for a slot (gpurun snapshots /root/repo when the box is acquired, not at submission).
#   tools/gpu_job.sh [--gpus N] [--timeout S] -- 'command run inside the frozen copy; write results to $OUT (= gpurun_out/)'
set -e
cd "$(dirname "$0")/.."
args=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do args+=("$1"); shift; done
shift
rm -rf .frozen/cur
mkdir -p .frozen/cur
tar -cf - --exclude=./.git --exclude=./gpurun_out --exclude=./.frozen --exclude='*.o' --exclude='*.ncu-rep' \
    --exclude=__pycache__ --exclude=.pytest_cache . | tar -xf - -C .frozen/cur
exec tools/gpurun_retry.sh "${args[@]}" -- "export DMV_NO_REBUILD=1 OUT=\$GRAFT_REPO_ROOT/gpurun_out; mkdir -p \$OUT; cd .frozen/cur && ( $* )"
