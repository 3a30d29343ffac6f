#!/usr/bin/env python
"""profiles/roofline_traffic.json from an `ncu --set full` capture of kernels INSIDE a bench.py step:
    python tools/ncu_traffic.py gpurun_out/bench_convtc.ncu-rep conv_tc_kernel "<how it was captured>" [more.ncu-rep kernel note ...]
Per kernel family: mean dram__bytes_read.sum + dram__bytes_write.sum per captured launch, the launch count, mean duration."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
args = sys.argv[1:]
for path, kern, note in zip(args[0::3], args[1::3], args[2::3]):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tscale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'second': 1e3, 's': 1e3}
    tot, n, ms, tens = 0.0, 0, 0.0, 0.0
    for r in rows[2:]:
        if kern not in r[ix['Kernel Name']]:
            continue
        b = 0.0
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            b += float(r[ix[m]]) * scale[units[ix[m]]]
        tot += b; n += 1
        ms += float(r[ix['gpu__time_duration.sum']]) * tscale[units[ix['gpu__time_duration.sum']]]
        tens += float(r[ix['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]) if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in ix else 0.0
    if n:
        out[kern] = dict(dram_bytes_per_launch=tot / n, launches_captured=n, mean_launch_ms_under_ncu=ms / n,
                         mean_tensor_pipe_active_pct=tens / n, note=f'mean over {n} launches captured with ncu --set full: {note} ({os.path.basename(path)})')
json.dump(out, open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
