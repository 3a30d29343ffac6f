#!/usr/bin/env python
"""Run one tcgen05 conv case per subprocess under different GG_TC_DBG switches; report time + outcome."""
import os, sys, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, time
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "ga-gan_b200"))
import torch
from torch_utils import custom_ops
sys.path.insert(0, os.path.join(%r, "tools"))
import tc_debug
t0 = time.time()
try:
    tc_debug.check(*eval(sys.argv[1]))
    print("OK   %%.2fs" %% (time.time() - t0))
except Exception as e:
    print("FAIL %%.2fs %%s" %% (time.time() - t0, str(e).splitlines()[0]))
''' % (ROOT, ROOT, ROOT)
cases = [('0', '(1,64,16,16,16,3)'), ('0', '(1,16,128,16,16,3)'), ('0', '(2,16,16,16,16,3)'), ('0', '(1,16,16,32,32,3)'),
         ('0', '(1,16,64,16,16,3)'), ('0', '(1,16,32,16,16,3)')]
for dbg, args in cases:
    env = dict(os.environ, GG_TC_DBG=dbg)
    t0 = time.time()
    r = subprocess.run([sys.executable, '-c', code, args], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    print(f'--- GG_TC_DBG={dbg} args={args} ({time.time()-t0:.1f}s)')
    for l in lines[-6:]:
        print('   ', l[:300])
