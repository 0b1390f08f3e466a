"""Canonical traversal of (okey, Dat, icfg).

Mirrors `/root/reference/src/obsiter.jl:29-59`: configs in index order, and
within a config the keys of `cfg.D` in the Dict's iteration order (`:43`).
Julia's `Dict{String}` order is hash order; here it is the insertion order of
`Dat.D` (E, F, V unless the caller chose otherwise).  Row numbering is derived
from this traversal only (`src/lsq_db.jl:237-250`), and the C-ABI receives the
resulting offsets explicitly, so any key order round-trips.

`tfor_observations` (`:81-105`) -- the threaded (config, okey) task loop -- has
no counterpart: the CUDA path takes the whole config list in one call.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Tuple


def observations(configs_or_db, Icfg: Iterable[int] = None) -> Iterator[Tuple[str, object, int]]:
    configs = getattr(configs_or_db, "configs", configs_or_db)
    idx: List[int] = list(range(len(configs))) if Icfg is None else list(Icfg)
    for icfg in idx:
        d = configs[icfg]
        for okey in list(d.D.keys()):
            yield okey, d, icfg
