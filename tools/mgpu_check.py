"""Hardware multi-GPU parity check (launch under torchrun with N >= 2 ranks, one GPU each).

Every rank assembles the rows of its own shard (configuration c -> rank c mod N) and then ALSO the whole set on its own
GPU.  Asserted:
  * the shard's design-matrix rows are BIT-identical to the same rows of the single-GPU matrix (the assembly has no
    collective and no atomics: any sharding gives the same bits);
  * the N-GPU fit (Gram blocks summed by the library's NCCL communicator) returns the same coefficients on every rank
    (bit-identical) and agrees with the single-GPU fit to max(1e-8, 100 cond eps), rel_rms and the per-configtype RMSE
    table to 1e-8.
  * the same holds for the TSQR exchange and for a context without a library communicator whose Gram blocks the host
    program sums itself (torch.distributed).
Prints one JSON line on rank 0.  Used by tests/test_multigpu.py and run by hand for profiles/."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import ipfitting_jl_b200 as ipf
    from ipfitting_jl_b200 import _capi, parallel, synth
    from ipfitting_jl_b200.lsq_db import LsqDB

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    torch.cuda.set_device(local)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                      # NCCL banners go to stderr
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = parallel.Communicator()
    ctx = _capi.Context(local)
    comm.attach(ctx)

    r0 = ipf.rnn("Si")
    desc = ipf.BondAngleDesc(f"exp(- (r/{r0} - 1.0))", ipf.CosCut(2.0 * r0 - 1.0, 2.0 * r0))
    basis = ipf.NBodyBasis(ipf.nbpolys(2, desc, 6) + ipf.nbpolys(3, desc, 4) + ipf.nbpolys(4, desc, 3))
    calc = synth.StillingerWeber()
    full = []
    for c in range(ncfg):
        d = synth.rattled_configs("Si", 2, 1, seed=1000 + c, calc=calc, min_dist=0.75, mode="julip",
                                  configtype="even" if c % 2 == 0 else "odd")
        full += d
    weights = {"default": {"E": 100.0, "F": 1.0, "V": 1.0}, "odd": {"E": 30.0, "F": 2.0, "V": 0.5}}

    # ---- single GPU: the whole set on this rank's device, no communicator
    ctx1 = _capi.Context(local)
    db1 = LsqDB("", basis, full, ctx=ctx1)
    Psi1 = db1.Psi.copy()
    rows1 = [dict(d.rows) for d in full]
    IP1, info1 = ipf.lsqfit(db1, weights={k: dict(v) for k, v in weights.items()}, E0=-4.0, Vref=ipf.OneBody(-4.0),
                            error_table=True)
    db1.delete()

    # ---- N GPUs: own shard, Gram all-reduce inside the library
    mine = list(range(rank, ncfg, world))
    shard = [ipf.Dat(full[c].at, full[c].configtype, E=full[c].D["E"], F=full[c].D["F"], V=full[c].D["V"]) for c in mine]
    dbN = LsqDB("", basis, shard, ctx=ctx)
    PsiN = dbN.Psi
    bit_identical = True
    for d, c in zip(shard, mine):
        for k, (r, n) in d.rows.items():
            r1, n1 = rows1[c][k]
            bit_identical &= bool(n == n1 and np.array_equal(PsiN[r:r + n].view(np.int64), Psi1[r1:r1 + n1].view(np.int64)))
    IPN, infoN = ipf.lsqfit(dbN, weights={k: dict(v) for k, v in weights.items()}, E0=-4.0, Vref=ipf.OneBody(-4.0),
                            error_table=True, comm=comm)
    # the TSQR exchange: every rank factors its own rows, one exchange of the [R_p | Q_p^T y] blocks
    IPT, infoT = ipf.lsqfit(dbN, weights={k: dict(v) for k, v in weights.items()}, E0=-4.0, Vref=ipf.OneBody(-4.0),
                            solver={"solver": "qr", "exchange": "tsqr"}, comm=comm)
    # the exchange left to the caller: a context WITHOUT a library communicator, Gram blocks summed by torch.distributed
    ctxX = _capi.Context(local)
    dbX = LsqDB("", basis, shard, ctx=ctxX)
    IPX, infoX = ipf.lsqfit(dbX, weights={k: dict(v) for k, v in weights.items()}, E0=-4.0, Vref=ipf.OneBody(-4.0),
                            comm=parallel.Communicator())
    dbX.delete()
    cN, c1 = infoN["c"], info1["c"]
    kappa = float(info1["cond_R"])
    dc = float(np.linalg.norm(cN - c1) / np.linalg.norm(c1))
    gathered = comm.allgather_object(cN.tobytes())
    same_on_all_ranks = all(g == gathered[0] for g in gathered)
    flags = comm.allgather_object(bool(bit_identical))
    drms = 0.0
    for ct in info1["errors"]["rmse"]:
        for ok in info1["errors"]["rmse"][ct]:
            a, b = info1["errors"]["rmse"][ct][ok], infoN["errors"]["rmse"][ct][ok]
            drms = max(drms, abs(a - b) / abs(a))
    out = {"check": "multi-GPU parity", "n_gpus": world, "configs": ncfg, "cols": len(basis), "rows": int(Psi1.shape[0]),
           "psi_shard_rows_bit_identical_on_all_ranks": bool(all(flags)), "coefficients_bit_identical_across_ranks": bool(same_on_all_ranks),
           "coeff_rel_diff_vs_1gpu": dc, "cond_R": kappa, "coeff_tolerance": max(1e-8, 100 * kappa * 2.2e-16),
           "rel_rms_1gpu": float(info1["rel_rms"]), "rel_rms_ngpu": float(infoN["rel_rms"]),
           "max_rel_diff_rmse_table": drms, "passes": infoN["solve_info"].get("passes"),
           "tsqr_coeff_rel_diff_vs_1gpu": float(np.linalg.norm(infoT["c"] - c1) / np.linalg.norm(c1)),
           "tsqr_rel_rms": float(infoT["rel_rms"]), "tsqr_local_passes": infoT["solve_info"].get("passes_local"),
           "external_exchange_coeff_rel_diff_vs_1gpu": float(np.linalg.norm(infoX["c"] - c1) / np.linalg.norm(c1)),
           "external_exchange_rel_rms": float(infoX["rel_rms"])}
    ok = (out["psi_shard_rows_bit_identical_on_all_ranks"] and out["coefficients_bit_identical_across_ranks"]
          and dc <= out["coeff_tolerance"] and abs(out["rel_rms_1gpu"] - out["rel_rms_ngpu"]) <= 1e-8 * out["rel_rms_1gpu"]
          and drms <= 1e-8 and out["tsqr_coeff_rel_diff_vs_1gpu"] <= out["coeff_tolerance"]
          and abs(out["tsqr_rel_rms"] - out["rel_rms_1gpu"]) <= 1e-8 * out["rel_rms_1gpu"]
          and out["external_exchange_coeff_rel_diff_vs_1gpu"] <= out["coeff_tolerance"]
          and abs(out["external_exchange_rel_rms"] - out["rel_rms_1gpu"]) <= 1e-8 * out["rel_rms_1gpu"])
    out["ok"] = bool(ok)
    if rank == 0:
        sys.stdout.flush()
        os.dup2(saved, 1)
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
