#!/bin/bash
# Multi-GPU measurements of one node (N = $1 GPUs): NCCL DDP parity test, weak / strong scaling of the headline workload,
# BASELINE configs[2] (ADA Affine+ 256^2) and configs[3] (GA population eval).  Output: gpurun_out/r2_n$1_*.json
N=${1:-2}
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "${ONLY_TRAIN:-0}" != "1" ]; then python -m pytest tests/test_gpu_multigpu.py -m gpu -q -s > gpurun_out/r2_n${N}_pytest.log 2>&1; echo "pytest multigpu rc=$?"; grep -E "DDP|passed|failed|skipped" gpurun_out/r2_n${N}_pytest.log | tail -3; fi
if [ "${SKIP_WEAK:-0}" != "1" ]; then $TR bench.py --gpus $N --steps 8 --warmup 3 --no-fast-mode > gpurun_out/r2_n${N}_train_weak.json 2> gpurun_out/r2_n${N}_train_weak.err; echo "weak rc=$?"; fi
$TR bench.py --gpus $N --global-batch 32 --steps 8 --warmup 3 --no-fast-mode > gpurun_out/r2_n${N}_train_strong.json 2> gpurun_out/r2_n${N}_train_strong.err; echo "strong rc=$?"
if [ "${ONLY_TRAIN:-0}" != "1" ]; then
$TR bench.py --gpus $N --workload ada --steps 8 --warmup 3 > gpurun_out/r2_n${N}_ada.json 2> gpurun_out/r2_n${N}_ada.err; echo "ada rc=$?"
$TR bench.py --gpus $N --workload ga --steps 4 --warmup 2 > gpurun_out/r2_n${N}_ga.json 2> gpurun_out/r2_n${N}_ga.err; echo "ga rc=$?"
fi
for f in train_weak train_strong ada ga; do python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2_n${N}_$f.json') if l.startswith('{')][-1])
    print('$f', 'N=', d['n_gpus'], 'value', round(d['value'], 2), d['unit'], 'e2e', round(d['e2e']['value'], 2), 'ms/step', round(d['ms_per_step'], 1), d['config'].get('individuals_per_sec'))
except Exception as e:
    print('$f', 'no result:', e)
PY
done
