#!/bin/bash
# One 8-GPU session: parity at P=8, then bench lines for four workloads (both exchange modes for the small one).
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/multi_gpu_check.py heisenberg_chain_16 heisenberg_square_4x4 heisenberg_chain_24_symm 2>&1 | grep -E "OK|FAIL|rror" | tee gpurun_out/multi$N.log
fmt='
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d["config"]["workload"], "N=",d["n_gpus"], d["config"]["exchange"], "ms/step", round(d["ms_per_step"],3), "Gstates/s", round(d["value"]/1e9,3), "Gterms/s", round(d["config"]["terms_per_s"]/1e9,2), "e2e ms", round(d["e2e"]["ms_per_step"],3), {k[:10]:round(v,3) for k,v in d["e2e"]["stages_ms"].items()}, "build_s", round(d["config"]["basis_build_s"],2), "states", d["config"]["basis_states"], "terms", d["config"]["off_diag_terms"])
'
for w in heisenberg_chain_24 heisenberg_chain_32_symm heisenberg_square_6x6 heisenberg_chain_36_symm; do
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --workload $w 2>&1 | grep "^{" | tee -a gpurun_out/bench_lines_$N.jsonl | python -c "$fmt" | tee -a gpurun_out/scale$N.log
done
DMV_EXCHANGE=0 timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 10 --workload heisenberg_chain_24 2>&1 | grep "^{" | python -c "$fmt" | tee -a gpurun_out/scale$N.log
