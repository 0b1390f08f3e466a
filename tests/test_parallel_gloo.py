"""N > 1 host logic on CPU: two gloo ranks shard the configurations (LPT), each holds the
design-matrix rows of its own shard, the staged solve protocol sums the Gram blocks with ONE
all-reduce per pass, and the result equals the single-process Householder solution.
The device calls are replaced by a numpy stand-in with the same four methods (test-only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyStagedSolver:
    """CPU stand-in of ipf_solve_begin/gram/factor/finish (same algorithm: iterated CholeskyQR)."""

    def __init__(self, A, y):
        self.A = np.array(A, dtype=np.float64, order="F")
        self.y = np.array(y, dtype=np.float64)
        self.n = A.shape[1]
        self.Rtot = np.eye(self.n)
        self.passes = 0
        self.done = None

    def gram(self):
        Ay = np.column_stack([self.A, self.y])
        self.G = Ay.T @ Ay
        return self.G, self.n + 1

    def factor(self):
        n = self.n
        G, z = self.G[:n, :n], self.G[:n, n]
        defect = np.linalg.norm(G - np.eye(n))
        final = self.passes >= 1 and defect <= 0.1
        L = np.linalg.cholesky(G)
        self.Rtot = L.T @ self.Rtot
        if final:
            w = np.linalg.solve(L, z)
            self.cc = np.linalg.solve(L.T, w)
            self.c = np.linalg.solve(self.Rtot, w)
            return True
        self.A = np.linalg.solve(L, self.A.T).T
        self.passes += 1
        return False

    def finish(self, want_R=True):
        r = self.y - self.A @ self.cc
        return self.c, self.Rtot, float(r @ r), float(self.y @ self.y)

    def info(self):
        return {"passes": self.passes + 1}


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ipfitting_jl_b200 as ipf
        from ipfitting_jl_b200 import parallel, synth
        from ipfitting_jl_b200.lsq import run_staged
        from ipfitting_jl_b200.lsq_db import _alloc_rows
        from ipfitting_jl_b200.lsqerrors import _finish_errors
        import ipf_oracle as O
        import lsq_oracle as LO
        from util import rows_for, si_desc

        sw = synth.StillingerWeber()
        configs = (synth.rattled_configs("Si", 1, 5, seed=1, calc=sw, configtype="small")
                   + synth.rattled_configs("Si", 2, 3, seed=2, calc=sw, configtype="big"))
        B = ipf.NBodyBasis(ipf.nbpolys(2, si_desc(2.0), 4) + ipf.nbpolys(3, si_desc(1.7), 2))
        weights = {"default": {"E": 30.0, "F": 1.0, "V": 0.5}}
        comm = parallel.Communicator()
        shards = parallel.shard_configs([parallel.config_cost(len(d)) for d in configs], world)
        mine = [configs[i] for i in shards[rank]]
        n = _alloc_rows(mine)
        r = rows_for(mine)
        Psi = O.assemble(B, mine, r["E"], r["F"], r["V"], n)
        A, y = LO.get_lsq_system(Psi, mine, weights, E0=-4.0)
        c, R, info = run_staged(NumpyStagedSolver(A, y), comm)
        # per-configtype error sums with config types that differ between ranks
        cts = sorted({d.configtype for d in mine})
        sse = np.arange(3.0 * len(cts)) + 1 + rank
        ssy = 2 * sse
        cnt = np.full(3 * len(cts), 10, dtype=np.int64)
        sse2, ssy2, cnt2, allcts = comm.allreduce_named_stats(cts, ["E", "F", "V"], sse, ssy, cnt)
        out = {"rank": rank, "c": c, "rel": info["rel_rms"], "shard": shards[rank], "allcts": allcts,
               "cnt": cnt2.tolist(), "local_cts": cts}
        if rank == 0:
            nall = _alloc_rows(configs)
            ra = rows_for(configs)
            Pall = O.assemble(B, configs, ra["E"], ra["F"], ra["V"], nall)
            cref, kappa, relref, _ = LO.lsqfit_qr(Pall, configs, weights, E0=-4.0)
            out.update(cref=cref, relref=relref, kappa=kappa)
        q.put(out)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_two_rank_gloo_solve_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        o = q.get(timeout=240)
        res[o["rank"]] = o
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = res[0], res[1]
    assert sorted(r0["shard"] + r1["shard"]) == list(range(8))
    assert np.array_equal(r0["c"], r1["c"])                      # every rank factors the same reduced Gram
    assert np.linalg.norm(r0["c"] - r0["cref"]) <= 1e-8 * np.linalg.norm(r0["cref"])
    assert abs(r0["rel"] - r0["relref"]) <= 1e-8 * r0["relref"]
    assert r0["allcts"] == r1["allcts"] and set(r0["allcts"]) == {"small", "big"}
    assert r0["cnt"] == r1["cnt"]
