#!/bin/bash
# One N-GPU session: parity on a small model, then bench lines for the BASELINE workloads.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/multi_gpu_check.py heisenberg_chain_16 heisenberg_square_4x4 2>&1 | grep -E "OK|FAIL|rror" | tee gpurun_out/multi$N.log
fmt='
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d["config"]["workload"], "N=",d["n_gpus"], d["config"]["exchange"], "ms/step", round(d["ms_per_step"],3), "Gstates/s", round(d["value"]/1e9,3), "Gterms/s", round(d["config"]["terms_per_s"]/1e9,2), "e2e ms", round(d["e2e"]["ms_per_step"],3), {k[:10]:round(v,3) for k,v in d["e2e"]["stages_ms"].items()}, "build_s", round(d["config"]["basis_build_s"],2))
'
for w in ${WORKLOADS:-heisenberg_chain_24 heisenberg_square_6x6 heisenberg_chain_36_symm heisenberg_chain_32_symm}; do
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --workload $w 2>&1 | grep "^{" | tee -a gpurun_out/bench_lines_$N.jsonl | python -c "$fmt" | tee -a gpurun_out/scale$N.log
done
