import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np
import ipfitting_jl_b200 as ipf
from ipfitting_jl_b200 import synth, _capi
from ipfitting_jl_b200.lsq_db import LsqDB
ctx=_capi.Context(0)
W = {"default": {"E": 100.0, "F": 1.0}}
r0 = ipf.rnn("Cu"); calc = synth.PairPotential(r0)
data = synth.rattled_configs("Cu", 3, 70, rmax=0.25 * r0, strain=0.0, seed=1, calc=calc, virial=False, mode="julip")
rr = np.linspace(0.9 * r0, calc.rcut, 200); phi, dphi = calc.phi(rr)
desc = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(6.0, 7.0))
for deg in (8,12,16,18):
    B = ipf.nbpolys(2, desc, deg); db = LsqDB("", B, data, ctx=ctx)
    IP, info = ipf.lsqfit(db, Vref=ipf.OneBody(0.0), weights=dict(W), error_table=True)
    c=info["c"]; x=desc.x(rr); s=(rr-6.0); fc=desc.cutoff(rr); dfc=np.where((rr>6.0)&(rr<7.0), -0.5*np.pi*np.sin(np.pi*s),0.0)
    b=np.arange(1,deg+1); P=(c[None,:]*x[:,None]**b[None,:]).sum(1); dP=(c[None,:]*b[None,:]*x[:,None]**b[None,:]).sum(1)*(-1.0/r0)
    print("fit2b deg",deg,"rmseE %.2e (5e-6) rmseF %.2e (2e-4) uE %.2e (5e-5) uF %.2e (1e-4)"%(info["errors"]["rmse"]["rand"]["E"],info["errors"]["rmse"]["rand"]["F"],np.abs(fc*P-0.5*phi).max(),np.abs(dfc*P+fc*dP-0.5*dphi).max()), "passes", info["solve_info"]["passes"], "cond %.1e"%info["cond_R"])
    db.delete()
r0 = ipf.rnn("Si"); calc = synth.StillingerWeber()
data = synth.rattled_configs("Si", 2, 100, rmax=0.33 * r0, strain=0.0, seed=1, calc=calc, virial=False, mode="julip", configtype="set")
desc = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(calc.rcut - 0.5, calc.rcut))
for d2,d3 in ((8,6),(12,10),(14,14)):
    B = ipf.nbpolys(2, desc, d2) + ipf.nbpolys(3, desc, d3); db = LsqDB("", B, data, ctx=ctx)
    IP, info = ipf.lsqfit(db, E0=0.0, weights=dict(W), error_table=True)
    print("fit3b",d2,d3,"relE %.2e (2e-5) relF %.2e (5e-3)"%(info["errors"]["relrmse"]["set"]["E"],info["errors"]["relrmse"]["set"]["F"]), "passes", info["solve_info"]["passes"], "cond %.1e"%info["cond_R"])
    db.delete()
