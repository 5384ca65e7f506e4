#!/usr/bin/env python
"""Ring-geometry sweep for the CG / laplace kernels (diagnostic knobs PHICUDA_RING_TY / PHICUDA_RING_NZC):
   python tools/sweep_ring.py 256 512x512x64 ...      prints us per CG iteration for every (TY, NZC) combination."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shapes = [a for a in sys.argv[1:] if a[0].isdigit()] or ['256', '512x512x64']
for shape in shapes:
    for ty in ('', '2', '4', '8', '16'):
        for nzc in ('', '1', '2', '4'):
            env = dict(os.environ)
            if ty:
                env['PHICUDA_RING_TY'] = ty
            if nzc:
                env['PHICUDA_RING_NZC'] = nzc
            out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'microbench.py'), shape, '--ring-only'] + [a for a in sys.argv[1:] if a.startswith('--')],
                                 env=env, capture_output=True, text=True)
            line = [l for l in out.stdout.splitlines() if l.startswith('n=')]
            print(f"TY={ty or 'auto':4s} NZC={nzc or 'auto':4s} {line[-1] if line else out.stderr[-300:]}", flush=True)
