#!/bin/bash
# One-GPU closing measurements of a round: GPU tests, the default bench line, the secondary workloads, the ncu launch list of the bench
# command, the cfg 5 sweep and the per-layer microbenchmarks.  Output: gpurun_out/r2_final_*
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2_final_pytest.log
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"
python bench.py --workload ada --no-cpu-baseline > gpurun_out/r2_final_ada.json 2> gpurun_out/r2_final_ada.err; echo "ada rc=$?"
python bench.py --workload ga --no-cpu-baseline > gpurun_out/r2_final_ga.json 2> gpurun_out/r2_final_ga.err; echo "ga rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_final_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fast-mode > gpurun_out/r2_final_ncu.log 2>&1; echo "ncu rc=$?"
python tools/sweep_cfg5.py --out gpurun_out/r2_cfg5_sweep > gpurun_out/r2_cfg5_sweep.log 2>&1; echo "sweep rc=$?"
python tools/microbench.py --out gpurun_out/r2_final_microbench.json > gpurun_out/r2_final_microbench.txt 2>&1; echo "microbench rc=$?"
python tools/phase_times.py 2>&1 | tail -2 | tee gpurun_out/r2_final_phase_times.txt
for f in bench ada ga; do python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2_final_$f.json') if l.startswith('{')][-1])
    print('$f', round(d['value'], 2), d['unit'], 'e2e', round(d['e2e']['value'], 2), 'ms/step', round(d['ms_per_step'], 1), d.get('clocks'), 'launches', d.get('gpu_launches'),
          'roofline', round(d.get('roofline', {}).get('achieved', 0), 1), round(d.get('roofline', {}).get('frac', 0), 3), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'fast', d.get('config', {}).get('fast_mode'))
except Exception as e:
    print('$f', 'no result:', e)
PY
done
