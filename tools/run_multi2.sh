mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for ex in 0 1; do DMV_EXCHANGE=$ex $TR --master-port 2951$ex tools/multi_gpu_check.py heisenberg_chain_16 heisenberg_square_4x4 heisenberg_chain_24_symm heisenberg_chain_20 2>&1 | grep -E "OK|FAIL|rror" | tee -a gpurun_out/multi2.log; done
for w in heisenberg_chain_24 heisenberg_chain_32_symm; do for ex in 0 1; do DMV_EXCHANGE=$ex $TR --master-port 2952$ex bench.py --gpus 2 --steps 10 --workload $w 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'], 'N=',d['n_gpus'], d['config']['exchange'], 'ms/step', round(d['ms_per_step'],3), 'Gstates/s', round(d['value']/1e9,3), 'e2e ms', round(d['e2e']['ms_per_step'],3), {k[:10]:round(v,3) for k,v in d['e2e']['stages_ms'].items()})
" | tee -a gpurun_out/scale2.log; done; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
