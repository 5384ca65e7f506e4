import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phiflow_b200 import _ops as ops
from microbench import timed
vbc = (('periodic', 'periodic'),) * 3
for shape in ((256, 1024, 512), (512, 512, 512), (1024, 256, 512)):
    cells = shape[0] * shape[1] * shape[2]
    dom = ops.Domain(shape, (1.0, 1.0, 1.0), 1, vbc=vbc)
    x = torch.randn(dom._shape(dom.cext), device='cuda'); y = torch.empty_like(x)
    for ty in (4, 2):
        os.environ['PHICUDA_RING_TY'] = str(ty)
        a = timed(lambda: ops.laplace(dom, vbc, x, out=y), 20); b = timed(lambda: ops.laplace_axpy(dom, vbc, x, -12345.0, out=y), 20); c = timed(lambda: ops.laplace_axpy(dom, vbc, x, -54321.0, out=y), 20)
        print(f"{shape} TY={ty}: laplace {a*1e3:.1f} us, read-only {b*1e3:.1f} us, TMA only {c*1e3:.1f} us", flush=True)
    del x, y, dom
