"""Per-worker cycle counts of site4b_kernel (build with IPF_NVCC_DEFS=-DIPF_4B_TIMING; development aid)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import _capi
ctx = _capi.Context(0)
basis = bench.make_basis(); configs = bench.make_configs(range(int(sys.argv[1]) if len(sys.argv) > 1 else 128), 1)
lib = ctx.lib
out = np.zeros(3 * 16)
for rep in range(2):
    db = ipf.LsqDB("", basis, configs, ctx=ctx); db.Psi_dev; ctx.synchronize() if hasattr(ctx, "synchronize") else None
    import torch; torch.cuda.synchronize()
    nw = lib.ipf_debug_4b_timing(out.ctypes.data_as(C.POINTER(C.c_double)), 1)
    db.delete()
t = out[:3 * nw].reshape(3, nw)
print("raw cycle sums: phase-1 end max %.4g  barrier max %.4g  end max %.4g" % (t[0].max(), t[1].max(), t[2].max()))
print("phase-1 end (cycles / CTA-sum, relative to max):", np.round(t[0] / t[0].max(), 3))
print("phase-2 duration relative to max:", np.round((t[2] - t[1]) / (t[2] - t[1]).max(), 3))
print("total relative to max:", np.round(t[2] / t[2].max(), 3), " mean/max %.3f" % (t[2].mean() / t[2].max()))
print("phase-1 share of the slowest worker: %.3f" % (t[1].max() / t[2].max()))
