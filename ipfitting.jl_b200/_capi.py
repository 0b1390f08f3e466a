"""ctypes binding of libipf_b200.so (include/ipf_b200.h).  There is NO CPU
fallback: if the CUDA library is missing or no B200 is visible, every compute
entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _build

IPF_OK = 0
IPF_MAX_VARS = 10

_ERRNAMES = {-1: "IPF_EINVAL", -2: "IPF_ECUDA", -3: "IPF_ENOMEM", -4: "IPF_ENAN_Y", -5: "IPF_ENAN_W",
             -6: "IPF_ENAN_PSI", -7: "IPF_ESINGULAR", -8: "IPF_ESTATE", -9: "IPF_ENCCL"}
# messages the reference raises for the three NaN checks (src/lsq.jl:261-262,291)
_REF_MESSAGES = {-4: "NaN detected in observations vector", -5: "NaN detected in weights vector",
                 -6: "discovered NaNs in LSQ system matrix"}


class IPFError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {msg}")
        self.code = code


class BlockSpec(C.Structure):
    _fields_ = [("N", C.c_int32), ("nb", C.c_int32), ("nmono", C.c_int32), ("transform", C.c_int32),
                ("a", C.c_double), ("r0", C.c_double), ("rc1", C.c_double), ("rc2", C.c_double),
                ("mono_b", C.POINTER(C.c_int32)), ("mono_exp", C.POINTER(C.c_uint8)),
                ("zc", C.c_int32), ("zs", C.c_int32 * 4), ("env_t", C.c_int32),
                ("env_rc1", C.c_double), ("env_rc2", C.c_double)]


_lib = None
_lock = threading.Lock()

_vp = C.c_void_p
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)

# name -> (restype, argtypes).  tests/test_abi.py checks this table against include/ipf_b200.h
SIGNATURES = {
    "ipf_init": (C.c_int32, [C.c_int32, C.POINTER(_vp)]),
    "ipf_destroy": (C.c_int32, [_vp]),
    "ipf_last_error": (C.c_char_p, [_vp]),
    "ipf_launch_count": (C.c_int64, [_vp]),
    "ipf_stream": (_vp, [_vp]),
    "ipf_set_force_generic": (C.c_int32, [_vp, C.c_int32]),
    "ipf_set_async": (C.c_int32, [_vp, C.c_int32]),
    "ipf_profile_enable": (C.c_int32, [_vp, C.c_int32]),
    "ipf_profile_reset": (C.c_int32, [_vp]),
    "ipf_profile_get": (C.c_int32, [_vp, C.c_char_p, _dp, _lp]),
    "ipf_comm_unique_id": (C.c_int32, [_vp, _bp]),
    "ipf_comm_init": (C.c_int32, [_vp, C.c_int32, C.c_int32, _bp]),
    "ipf_comm_destroy": (C.c_int32, [_vp]),
    "ipf_comm_pause": (C.c_int32, [_vp, C.c_int32]),
    "ipf_comm_size": (C.c_int32, [_vp]),
    "ipf_comm_rank": (C.c_int32, [_vp]),
    "ipf_comm_allreduce_f64": (C.c_int32, [_vp, _dp, C.c_int64]),
    "ipf_basis_create": (C.c_int32, [_vp, C.c_int32, C.POINTER(BlockSpec), C.POINTER(_vp)]),
    "ipf_basis_destroy": (C.c_int32, [_vp, _vp]),
    "ipf_basis_length": (C.c_int64, [_vp]),
    "ipf_matrix_create": (C.c_int32, [_vp, C.c_int64, C.c_int64, C.POINTER(_vp)]),
    "ipf_matrix_destroy": (C.c_int32, [_vp, _vp]),
    "ipf_matrix_to_host": (C.c_int32, [_vp, _vp, _dp, C.c_int64]),
    "ipf_matrix_from_host": (C.c_int32, [_vp, _vp, _dp, C.c_int64]),
    "ipf_matrix_rows": (C.c_int64, [_vp]),
    "ipf_matrix_cols": (C.c_int64, [_vp]),
    "ipf_matrix_ld": (C.c_int64, [_vp]),
    "ipf_matrix_device_ptr": (_vp, [_vp]),
    "ipf_neighbours": (C.c_int32, [_vp, C.c_int32, _dp, _dp, _bp, C.c_double, C.c_int64, _lp, _ip, _ip, _ip, _dp]),
    "ipf_assemble": (C.c_int32, [_vp, _vp, C.c_int64, _lp, _dp, _dp, _bp, _ip, _lp, _lp, _lp, _vp]),
    "ipf_configs_create": (C.c_int32, [_vp, C.c_int64, _lp, _dp, _dp, _bp, _ip, _lp, _lp, _lp, C.POINTER(_vp)]),
    "ipf_configs_destroy": (C.c_int32, [_vp, _vp]),
    "ipf_assemble_configs": (C.c_int32, [_vp, _vp, _vp, _vp]),
    "ipf_solve_begin_dev": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, C.c_int64, _vp, C.c_int64, _vp, C.POINTER(_vp)]),
    "ipf_solve_begin": (C.c_int32, [_vp, _vp, _dp, _dp, _lp, C.c_int64, _lp, C.c_int64, _dp, C.POINTER(_vp)]),
    "ipf_solve_precon": (C.c_int32, [_vp, _vp, C.c_int64, _lp, _ip, _lp, _dp, C.c_int64, C.c_int32]),
    "ipf_solve_shape": (C.c_int32, [_vp, _lp, _lp]),
    "ipf_solve_system_to_host": (C.c_int32, [_vp, _vp, _dp, C.c_int64, _dp]),
    "ipf_solve_gram": (C.c_int32, [_vp, _vp, C.POINTER(_vp), _lp]),
    "ipf_solve_factor": (C.c_int32, [_vp, _vp, _ip]),
    "ipf_solve_finish": (C.c_int32, [_vp, _vp, _dp, _dp, _dp, _dp]),
    "ipf_solve_append_rows": (C.c_int32, [_vp, _vp, C.c_int64, _dp, C.c_int64, _dp]),
    "ipf_solve_qty": (C.c_int32, [_vp, _vp, _dp]),
    "ipf_solve_residual": (C.c_int32, [_vp, _vp, _dp, _dp, _dp]),
    "ipf_solve_info": (C.c_int32, [_vp, _ip, _dp, _dp, _dp, _dp]),
    "ipf_solve_cond": (C.c_int32, [_vp, _vp, _dp]),
    "ipf_solve_set_csne": (C.c_int32, [_vp, C.c_int32]),
    "ipf_solve_destroy": (C.c_int32, [_vp, _vp]),
    "ipf_solve": (C.c_int32, [_vp, _vp, _dp, _dp, _lp, C.c_int64, _lp, C.c_int64, _dp, _dp, _dp, _dp]),
    "ipf_predict": (C.c_int32, [_vp, _vp, _dp, _dp]),
    "ipf_residual_stats": (C.c_int32, [_vp, _vp, _dp, _dp, _dp, _dp, _ip, C.c_int32, _dp, _dp, _lp]),
}


def library_path() -> str:
    return _build.LIB


def load():
    """Load libipf_b200.so (does not touch the GPU)."""
    global _lib
    with _lock:
        if _lib is None:
            path = library_path()
            if not os.path.exists(path):
                raise IPFError(-2, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(there is no CPU fallback)")
            L = C.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


class Context:
    """One ipf_ctx = one GPU."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _vp()
        rc = self.lib.ipf_init(int(device), C.byref(h))
        self.h = h
        self.device = int(device)
        if rc != IPF_OK:
            msg = self.lib.ipf_last_error(h).decode() if h else "ipf_init failed"
            if h:
                self.lib.ipf_destroy(h)
            self.h = None
            raise IPFError(rc, msg + " (a CUDA device is required; there is no CPU fallback)")

    def check(self, rc: int):
        if rc != IPF_OK:
            msg = self.lib.ipf_last_error(self.h).decode()
            raise IPFError(rc, _REF_MESSAGES.get(rc, msg) if rc in _REF_MESSAGES else msg)

    def force_generic(self, on=True):
        """Kernel-path selection (tests / profiling): False/0 automatic, True/1 table-driven kernels only,
        2 = 3-body block through the HBM-scratch + gather path instead of the configuration-resident kernel."""
        self.lib.ipf_set_force_generic(self.h, int(on))

    def set_async(self, on: bool = True):
        """`ipf_assemble*` return without the final stream synchronisation (later calls are stream-ordered)."""
        self.lib.ipf_set_async(self.h, 1 if on else 0)

    def profile(self, on: bool = True):
        self.lib.ipf_profile_enable(self.h, 1 if on else 0)

    def profile_reset(self):
        self.lib.ipf_profile_reset(self.h)

    def profile_get(self, name: str):
        ms, cnt = C.c_double(), C.c_int64()
        self.lib.ipf_profile_get(self.h, name.encode(), C.byref(ms), C.byref(cnt))
        return ms.value, cnt.value

    @property
    def stream(self) -> int:
        return int(self.lib.ipf_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(self.lib.ipf_launch_count(self.h))

    # ---- multi-GPU communicator inside the library (include/ipf_b200.h: ipf_comm_*)
    def comm_unique_id(self) -> bytes:
        buf = np.zeros(128, dtype=np.uint8)
        self.check(self.lib.ipf_comm_unique_id(self.h, ptr(buf, C.c_uint8)))
        return buf.tobytes()

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = np.frombuffer(unique_id, dtype=np.uint8).copy()
        if buf.size != 128:
            raise ValueError("the NCCL unique id has 128 bytes")
        self.check(self.lib.ipf_comm_init(self.h, int(nranks), int(rank), ptr(buf, C.c_uint8)))

    def comm_pause(self, paused: bool):
        self.check(self.lib.ipf_comm_pause(self.h, 1 if paused else 0))

    def comm_size(self) -> int:
        return int(self.lib.ipf_comm_size(self.h))

    def comm_allreduce(self, x) -> np.ndarray:
        """Sum of a small host vector over the ranks of the attached communicator."""
        a = np.ascontiguousarray(x, dtype=np.float64).copy()
        self.check(self.lib.ipf_comm_allreduce_f64(self.h, ptr(a, C.c_double), a.size))
        return a

    def close(self):
        if self.h:
            self.lib.ipf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: int = None) -> Context:
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DeviceBasis:
    def __init__(self, ctx: Context, basis):
        self.ctx = ctx
        blocks = basis.blocks
        arr = (BlockSpec * len(blocks))()
        keep = []
        for k, b in enumerate(blocks):
            kind, a, r0, rc1, rc2 = b.desc.params()
            mb = np.ascontiguousarray(b.mono_b, dtype=np.int32)
            me = np.ascontiguousarray(b.mono_exp, dtype=np.uint8)
            keep += [mb, me]
            zs = (C.c_int32 * 4)(*([int(z) for z in getattr(b, "zs", ())] + [0] * 4)[:4])
            arr[k] = BlockSpec(b.N, b.nb, len(mb), kind, a, r0, rc1, rc2, ptr(mb, C.c_int32), ptr(me, C.c_uint8),
                               int(getattr(b, "zc", 0)), zs, int(getattr(b, "env_t", 0)),
                               float(getattr(b, "env_rc", (0.0, 0.0))[0]), float(getattr(b, "env_rc", (0.0, 0.0))[1]))
        h = _vp()
        ctx.check(ctx.lib.ipf_basis_create(ctx.h, len(blocks), arr, C.byref(h)))
        self.h = h

    def __len__(self):
        return int(self.ctx.lib.ipf_basis_length(self.h))

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.ipf_basis_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMatrix:
    """Device-resident column-major float64 matrix (db.Psi)."""

    def __init__(self, ctx: Context, nrows: int, ncols: int):
        self.ctx = ctx
        h = _vp()
        ctx.check(ctx.lib.ipf_matrix_create(ctx.h, int(nrows), int(ncols), C.byref(h)))
        self.h = h
        self.shape = (int(nrows), int(ncols))

    @property
    def ld(self):
        return int(self.ctx.lib.ipf_matrix_ld(self.h))

    @property
    def device_ptr(self) -> int:
        return int(self.ctx.lib.ipf_matrix_device_ptr(self.h) or 0)

    def to_host(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.float64, order="F")
        if out.size:
            self.ctx.check(self.ctx.lib.ipf_matrix_to_host(self.ctx.h, self.h, ptr(out, C.c_double), max(1, self.shape[0])))
        return out

    def from_host(self, A: np.ndarray):
        A = np.asfortranarray(A, dtype=np.float64)
        if A.shape != self.shape:
            raise ValueError("shape mismatch")
        if A.size:
            self.ctx.check(self.ctx.lib.ipf_matrix_from_host(self.ctx.h, self.h, ptr(A, C.c_double), max(1, self.shape[0])))

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.ipf_matrix_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def neighbours(ctx: Context, X, cell, pbc, rcut):
    X = np.ascontiguousarray(X, dtype=np.float64)
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    pbc = np.ascontiguousarray(pbc, dtype=np.uint8)
    n = C.c_int64(0)
    ctx.check(ctx.lib.ipf_neighbours(ctx.h, len(X), ptr(X, C.c_double), ptr(cell, C.c_double), ptr(pbc, C.c_uint8),
                                     float(rcut), 0, C.byref(n), None, None, None, None))
    m = n.value
    I = np.empty(m, dtype=np.int32); J = np.empty(m, dtype=np.int32)
    S = np.empty((m, 3), dtype=np.int32); R = np.empty((m, 3), dtype=np.float64)
    ctx.check(ctx.lib.ipf_neighbours(ctx.h, len(X), ptr(X, C.c_double), ptr(cell, C.c_double), ptr(pbc, C.c_uint8),
                                     float(rcut), m, C.byref(n), ptr(I, C.c_int32), ptr(J, C.c_int32),
                                     ptr(S, C.c_int32), ptr(R, C.c_double)))
    return I, J, S, R
