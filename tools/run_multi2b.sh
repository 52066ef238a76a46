mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29514 tools/multi_gpu_check.py heisenberg_square_4x4 heisenberg_chain_16 heisenberg_kagome_12_symm 2>&1 | grep -E "OK|FAIL|rror" | tee -a gpurun_out/multi${N}b.log
for w in heisenberg_chain_32_symm heisenberg_square_6x6; do for ex in 1 2; do DMV_EXCHANGE=$ex timeout 300 $TR --master-port 2952$ex bench.py --gpus $N --steps 5 --workload $w 2>&1 | grep "^{" | tee -a gpurun_out/bench_lines_${N}b.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'], 'N=',d['n_gpus'], d['config']['exchange'], 'ms/step', round(d['ms_per_step'],3), 'Gstates/s', round(d['value']/1e9,3), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k[:10]:round(v,3) for k,v in d['e2e']['stages_ms'].items()})
" | tee -a gpurun_out/scale${N}b.log; done; done
