"""`IPFitting.Data`: reading training sets (`/root/reference/src/data.jl`).

  read_xyz  :123-243  extended-xyz file -> Vector{Dat}; keys `dft_energy` / `dft_force` / `dft_virial`
                      (case-insensitive, `energy_key=` ... overrides), `config_type`, `include` / `exclude`
  load_data :248-256  dispatch on the file extension

The reference hands the parsing to `ase.io.read` through PyCall; `ase` is not part of this image, so the
extended-xyz frame format is parsed here directly (line 1: atom count; line 2: key=value pairs with `Lattice`,
`Properties`, `pbc`; then one row per atom with the columns `Properties` declares).  Nine-element values are 3 x 3
matrices stored column by column (ASE's convention; the cell's ROWS are the lattice vectors).  `write_xyz` is the
inverse (used by the tests; the reference has no writer)."""
from __future__ import annotations

import re
import shlex
from typing import Iterable, List, Optional, Sequence

import numpy as np

from .prototypes import Atoms, Dat

_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
            "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu "
            "Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()
_Z_OF = {s: z for z, s in enumerate(_SYMBOLS)}


def _parse_value(v: str):
    t = v.strip()
    parts = t.replace(",", " ").split()
    if len(parts) > 1:
        try:
            return np.array([float(p) for p in parts])
        except ValueError:
            if all(p in ("T", "F", "True", "False") for p in parts):
                return np.array([p in ("T", "True") for p in parts])
            return t
    if t in ("T", "True"):
        return True
    if t in ("F", "False"):
        return False
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(t)
    except ValueError:
        return t


def _parse_info(line: str) -> dict:
    lex = shlex.shlex(line, posix=True)
    lex.whitespace_split = True
    lex.commenters = ""
    info = {}
    for tok in lex:
        if "=" not in tok:
            info[tok] = True
            continue
        k, v = tok.split("=", 1)
        info[k] = v if k == "Properties" else _parse_value(v)
    return info


def _frames(fname: str):
    with open(fname, "r") as f:
        while True:
            head = f.readline()
            if not head:
                return
            if not head.strip():
                continue
            nat = int(head.split()[0])
            info = _parse_info(f.readline().rstrip("\n"))
            rows = [f.readline().split() for _ in range(nat)]
            yield nat, info, rows


def _select(frames: Sequence, index):
    if index in (":", None):
        return list(frames)
    if isinstance(index, int):
        return [list(frames)[index]]
    if isinstance(index, slice):
        return list(frames)[index]
    m = re.fullmatch(r"(-?\d*):(-?\d*)(?::(-?\d+))?", str(index))
    if not m:
        raise ValueError(f"unsupported index {index!r}")
    a, b, c = (int(x) if x not in (None, "") else None for x in m.groups())
    return list(frames)[slice(a, b, c)]


def _find(d: dict, key: str):
    for k in d:
        if k.lower() == key.lower():        # `lowercase(key) == lowercase(energy_key)`, src/data.jl:53-80
            return d[k]
    return None


def read_xyz(fname: str, *, energy_key: str = "dft_energy", force_key: str = "dft_force", virial_key: str = "dft_virial",
             auto: bool = False, verbose: bool = True, index=":", include: Optional[Iterable[str]] = None,
             exclude: Optional[Iterable[str]] = None) -> List[Dat]:
    """`read_xyz(fname; energy_key, force_key, virial_key, auto, verbose, index, include, exclude)` (`src/data.jl:129-243`)."""
    frames = _select(list(_frames(fname)), index)
    if auto:                                # closest key by edit-distance similarity (`compare(..., Levenshtein())`)
        import difflib
        infos = {k for _, info, _ in frames for k in info}
        arrays = {p for _, info, _ in frames for p in str(info.get("Properties", "")).split(":")[0::3]}
        best = lambda want, pool: max(pool, key=lambda k: difflib.SequenceMatcher(None, want, k).ratio()) if pool else want
        energy_key, force_key, virial_key = best(energy_key, infos), best(force_key, arrays), best(virial_key, infos)
    if verbose:
        print(f'Keys used: E => "{energy_key}", F => "{force_key}", V => "{virial_key}"')
    data: List[Dat] = []
    for nat, info, rows in frames:
        ct = _find(info, "config_type")
        if ct is None:
            if verbose:
                print(f"{len(data) + 1} has no config_type")
            ct = "nothing"
        ct = str(ct)
        if exclude is not None and ct in exclude:
            continue
        if include is not None and ct not in include:
            continue
        # per-atom columns
        props = str(info.get("Properties", "species:S:1:pos:R:3")).split(":")
        cols, c0 = {}, 0
        for name, typ, cnt in zip(props[0::3], props[1::3], props[2::3]):
            cols[name] = (c0, c0 + int(cnt), typ)
            c0 += int(cnt)
        arr = {}
        for name, (a, b, typ) in cols.items():
            if typ in ("R", "I"):
                arr[name] = np.array([[float(x) for x in r[a:b]] for r in rows], dtype=np.float64).reshape(nat, b - a)
            else:
                arr[name] = [r[a] for r in rows]
        species = arr.get("species") or arr.get("Z")
        Z = np.array([_Z_OF.get(s, 0) if not str(s).isdigit() else int(s) for s in (species or ["X"] * nat)], dtype=np.int32)
        lat = info.get("Lattice")
        if lat is None:
            cell, pbc = np.zeros((3, 3)), (False, False, False)
        else:
            cell = np.asarray(lat, dtype=np.float64).reshape(3, 3)      # a, b, c written one after the other = rows
            pbc = info.get("pbc", np.array([True, True, True]))
        at = Atoms(arr["pos"], cell, tuple(bool(x) for x in np.atleast_1d(pbc)), Z)
        E = _find(info, energy_key)
        F = _find({k: v for k, v in arr.items() if k not in ("species", "pos", "Z")}, force_key)
        V = _find(info, virial_key)
        if V is not None:
            V = np.asarray(V, dtype=np.float64).reshape(3, 3, order="F")
        data.append(Dat(at, ct, E=None if E is None else float(E), F=F, V=V))
    if verbose:
        cts = {}
        for d in data:
            t = cts.setdefault(d.configtype, [0, 0, 0, 0, 0])
            t[0] += 1; t[1] += len(d.at)
            t[2] += len(d.D.get("E", ())); t[3] += len(d.D.get("F", ())); t[4] += 9 if "V" in d.D else 0
        print("config_type  #cfgs  #envs  #E  #F  #V")
        for ct, t in cts.items():
            print(f"{ct:12s} " + "  ".join(str(x) for x in t))
    return data


def write_xyz(fname: str, data: Sequence[Dat], *, energy_key: str = "dft_energy", force_key: str = "dft_force",
              virial_key: str = "dft_virial") -> None:
    """Extended-xyz writer (17 significant digits: a read_xyz round trip is exact)."""
    from .datatypes import devec_obs
    g = lambda x: repr(float(x))
    with open(fname, "w") as f:
        for d in data:
            at = d.at
            F = devec_obs("F", d.D["F"]) if "F" in d.D else None
            props = "species:S:1:pos:R:3" + (f":{force_key}:R:3" if F is not None else "")
            items = ['Lattice="' + " ".join(g(x) for x in at.cell.reshape(-1)) + '"', f"Properties={props}",
                     'pbc="' + " ".join("T" if p else "F" for p in at.pbc) + '"', f"config_type={d.configtype}"]
            if "E" in d.D:
                items.append(f"{energy_key}={g(d.D['E'][0])}")
            if "V" in d.D:
                V = devec_obs("V", d.D["V"])
                items.append(f'{virial_key}="' + " ".join(g(x) for x in np.asarray(V).reshape(-1, order="F")) + '"')
            f.write(f"{len(at)}\n" + " ".join(items) + "\n")
            for i in range(len(at)):
                row = [_SYMBOLS[int(at.Z[i])] if 0 <= int(at.Z[i]) < len(_SYMBOLS) else "X"] + [g(x) for x in at.X[i]]
                if F is not None:
                    row += [g(x) for x in np.asarray(F)[i]]
                f.write(" ".join(row) + "\n")


def load_data(fname: str, **kwargs) -> List[Dat]:
    """`load_data(fname; kwargs...)` (`src/data.jl:248-256`): `.xyz` files; the JLD branch has no counterpart."""
    if fname.endswith(".xyz"):
        return read_xyz(fname, **kwargs)
    raise ValueError(f"unknown file format {fname[-4:]}")
