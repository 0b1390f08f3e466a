"""Hardware multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): tools/mgpu_check.py under torchrun asserts that
the N-GPU design-matrix slices are bit-identical to the single-GPU matrix and that the N-GPU fit (NCCL all-reduce of the
Gram blocks inside the library) reproduces the single-GPU coefficients, rel_rms and RMSE table."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpus_match_one(ctx):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 CUDA devices")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py"), "12"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"], out


def test_two_contexts_on_two_devices_in_one_process(ctx):
    """`distinct ctxs are independent` (include/ipf_b200.h): a second context on ANOTHER device of the same process runs
    every kernel family (the > 48 KB shared-memory opt-ins are per device) and returns bit-identical results."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 CUDA devices")
    import ipfitting_jl_b200 as ipf
    from ipfitting_jl_b200 import _capi, synth
    from ipfitting_jl_b200.lsq_db import LsqDB
    from util import si_desc
    B = ipf.nbpolys(2, si_desc(), 8) + ipf.nbpolys(3, si_desc(2.0), 6) + ipf.nbpolys(4, si_desc(1.8), 4)
    configs = synth.rattled_configs("Si", 2, 4, seed=13, calc=synth.StillingerWeber(), min_dist=0.75)
    ctx1 = _capi.Context(1)
    try:
        out = []
        for c in (ctx, ctx1, ctx):
            db = LsqDB("", B, configs, ctx=c)
            IP, info = ipf.lsqfit(db, weights={"default": {"E": 30.0, "F": 1.0, "V": 1.0}}, E0=-4.0)
            out.append((db.Psi.copy(), info["c"].copy()))
            db.delete()
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
        assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])
    finally:
        ctx1.close()
