import numpy as np, torch, sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np as O
from phiflow_b200 import _ops as ops, _lib
vbc=((0.0,0.0),(0.0,0.0)); d=2
sbc=O.uniform_bc(2,0.0)
rng=np.random.default_rng(3); res=(37,22); dx=(0.5,0.25)
lower=(0.0,0.0); upper=tuple(r*h for r,h in zip(res,dx))
shapes=O.staggered_shapes(res,vbc)
v=[c*np.float32(1.7) for c in [rng.standard_normal(s).astype(np.float32) for s in shapes]]
s=rng.standard_normal(res).astype(np.float32)
dom=ops.Domain(res,dx,1,vbc=vbc)
dv=dom.faces_from_numpy(v,vbc); ds=dom.centered_from_numpy(s)
out=dom.alloc_centered(); tmp=dom.alloc_centered()
_lib.check(_lib.load().phicuda_mac_cormack_centered_f32(C.byref(dom.grid), C.byref(ops.make_vbc(vbc,2)), ops._f3(dv), C.byref(ops.make_bc(sbc)),
    ops._ptr(ds), ops._ptr(out), ops._ptr(tmp), C.c_float(0.8), C.c_float(1.0), ops._stream()))
torch.cuda.synchronize()
sl=dom.centered_to_numpy(ops.advect_centered(dom,vbc,dv,sbc,ds,0.8))
t=dom.centered_to_numpy(tmp); got=dom.centered_to_numpy(out)
ref=O.mac_cormack_centered(s,sbc,v,vbc,lower,upper,0.8)
print('max|tmp-sl|',np.abs(t-sl).max(),'bad',(np.abs(got-ref)>1e-4).sum())
# emulate pass 2 on CPU from GPU tmp
v0=O._velocity_at_centers(v,res,vbc); pts=O.points_of(lower,upper,res)
cf=O.to_index_space(pts+v0*np.float32(0.8),lower,upper,res)
bw=O.grid_sample(t,cf,sbc)
print('fwd tmp[0,18:20]',t[0,18:20], 'cpu bwd(0,16) from gpu tmp', bw[0,16])
print('raw tmp tensor row y=19, x0..3', tmp[0,19,:4].cpu().numpy(), 'y=18', tmp[0,18,:4].cpu().numpy())
