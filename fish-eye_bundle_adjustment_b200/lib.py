"""ctypes binding of the C ABI in ``include/feba.h`` (``libfeba.so``, built by ``build.py``).

This is the same binding a MATLAB MEX gateway makes (``INTEGRATION.md``): plain pointers and
sizes, one opaque handle.  There is no fallback: when the shared object is missing, or no CUDA
device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .problem import Problem

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfeba.so")

FEBA_OK, FEBA_ERR_INVALID, FEBA_ERR_CUDA, FEBA_ERR_NUMERIC, FEBA_ERR_STATE = 0, 1, 2, 3, 4

# every symbol include/feba.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = (
    "feba_create", "feba_destroy", "feba_last_error", "feba_set_stream", "feba_num_unknowns", "feba_num_obs",
    "feba_set_xhat", "feba_get_xhat", "feba_iterate", "feba_iterate_assemble", "feba_reduced_dev",
    "feba_iterate_solve", "feba_get_delta", "feba_residuals", "feba_solve", "feba_last_timing",
    "feba_launch_count", "feba_sparse_info", "feba_debug_reduced", "feba_cov_prepare", "feba_cov_diag", "feba_cov_block", "feba_iterate_async", "feba_iterate_solve_async", "feba_sync",
    "feba_dist_unique_id", "feba_dist_init", "feba_reduced_pack", "feba_reduced_unpack",
    "feba_create_shard", "feba_last_timing_ex", "feba_plan_info",
    "feba_batch_create", "feba_batch_iterate_async", "feba_batch_destroy",
    "feba_set_xhat_owned", "feba_get_xhat_owned", "feba_num_owned_ties",
)


DIST_ID_BYTES = 128


def dist_unique_id() -> bytes:
    """feba_dist_unique_id: the token one rank creates and every rank passes to ``Handle.dist_init``."""
    lib = load()
    buf = C.create_string_buffer(DIST_ID_BYTES)
    if lib.feba_dist_unique_id(buf, DIST_ID_BYTES) != 0:
        raise FebaError(3, (lib.feba_last_error(None) or b"").decode())
    return buf.raw


class FebaSettings(C.Structure):
    _fields_ = [("estimate_eop", C.c_int32 * 6), ("estimate_xp", C.c_int32), ("estimate_yp", C.c_int32),
                ("estimate_c", C.c_int32), ("estimate_radial", C.c_int32), ("num_radial", C.c_int32),
                ("estimate_decent", C.c_int32), ("inner_constraints", C.c_int32), ("type", C.c_int32),
                ("iteration_cap", C.c_int32), ("plan", C.c_int32), ("sigma_x", C.c_double),
                ("sigma_y", C.c_double), ("threshold", C.c_double)]


_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class FebaProblem(C.Structure):
    _fields_ = [("n_obs", C.c_int64), ("n_img", C.c_int32), ("n_cam", C.c_int32), ("n_pts", C.c_int32),
                ("n_tie", C.c_int32), ("obs_x", _pd), ("obs_y", _pd), ("obs_img", _pi), ("obs_pt", _pi),
                ("img_cam", _pi), ("eop0", _pd), ("iop0", _pd), ("cam_box", _pd), ("xyz0", _pd),
                ("pt_tie", _pi), ("settings", FebaSettings)]


class FebaError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"feba error {code}: {text}")
        self.code = code
        self.text = text


_lib = None


def load() -> C.CDLL:
    """Load libfeba.so; raises when it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FebaError(FEBA_ERR_CUDA, f"{LIB_PATH} is missing: run __graft_entry__.build() "
                                       "(python fish-eye_bundle_adjustment_b200/build.py)")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.feba_create.argtypes = [C.POINTER(FebaProblem), C.POINTER(H)]
    lib.feba_create_shard.argtypes = [C.POINTER(FebaProblem), C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(H)]
    lib.feba_last_timing_ex.argtypes = [H, _pd]
    lib.feba_plan_info.argtypes = [H, C.POINTER(C.c_int32), _pd]
    lib.feba_batch_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p)]
    lib.feba_batch_iterate_async.argtypes = [C.c_void_p]
    lib.feba_batch_destroy.argtypes = [C.c_void_p]
    lib.feba_batch_destroy.restype = None
    lib.feba_destroy.argtypes = [H]
    lib.feba_destroy.restype = None
    lib.feba_last_error.argtypes = [H]
    lib.feba_last_error.restype = C.c_char_p
    lib.feba_set_stream.argtypes = [H, C.c_void_p]
    lib.feba_num_unknowns.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.feba_num_obs.argtypes = [H]
    lib.feba_num_obs.restype = C.c_int64
    lib.feba_set_xhat.argtypes = [H, _pd, C.c_size_t]
    lib.feba_get_xhat.argtypes = [H, _pd, C.c_size_t]
    lib.feba_get_delta.argtypes = [H, _pd, C.c_size_t]
    lib.feba_set_xhat_owned.argtypes = [H, _pd, C.c_size_t]
    lib.feba_get_xhat_owned.argtypes = [H, _pd, C.c_size_t]
    lib.feba_num_owned_ties.argtypes = [H]
    lib.feba_num_owned_ties.restype = C.c_int64
    lib.feba_iterate.argtypes = [H, _pd]
    lib.feba_iterate_assemble.argtypes = [H]
    lib.feba_reduced_dev.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.feba_iterate_solve.argtypes = [H, _pd, _pd]
    lib.feba_reduced_pack.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.feba_reduced_unpack.argtypes = [H]
    lib.feba_residuals.argtypes = [H, _pd, _pd, _pd]
    lib.feba_solve.argtypes = [H, _pi, _pd, C.c_size_t]
    lib.feba_last_timing.argtypes = [H, _pd]
    lib.feba_launch_count.argtypes = [H]
    lib.feba_sparse_info.argtypes = [H, C.POINTER(C.c_int32)]
    lib.feba_launch_count.restype = C.c_int64
    lib.feba_debug_reduced.argtypes = [H, _pd, _pd]
    lib.feba_cov_prepare.argtypes = [H]
    lib.feba_cov_diag.argtypes = [H, _pd, C.c_size_t]
    lib.feba_cov_block.argtypes = [H, C.POINTER(C.c_int64), C.c_int32, _pd]
    lib.feba_iterate_async.argtypes = [H]
    lib.feba_iterate_solve_async.argtypes = [H]
    lib.feba_sync.argtypes = [H, _pd]
    lib.feba_dist_unique_id.argtypes = [C.c_void_p, C.c_size_t]
    lib.feba_dist_init.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]
    _lib = lib
    return lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_pd)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(_pi)


def settings_struct(s, plan: int = 0) -> FebaSettings:
    """``data.settings`` -> ``feba_settings`` (include/feba.h).  ``plan``: row order of the reduced system
    (0 automatic, -1 Buildxhat order / dense, 1 force nested dissection)."""
    fs = FebaSettings()
    fs.plan = int(plan)
    for q, f in enumerate(s.eop_flags):
        fs.estimate_eop[q] = int(f)
    fs.estimate_xp, fs.estimate_yp, fs.estimate_c = int(s.Estimate_xp), int(s.Estimate_yp), int(s.Estimate_c)
    fs.estimate_radial, fs.num_radial = int(s.Estimate_radial), int(s.NK)
    fs.estimate_decent, fs.inner_constraints = int(s.Estimate_decent), int(s.Inner_Constraints)
    fs.type = int(s.typeint)
    fs.iteration_cap = int(s.Iteration_Cap)
    fs.sigma_x, fs.sigma_y, fs.threshold = float(s.sigma_x), float(s.sigma_y), float(s.threshold)
    return fs


class Handle:
    """One adjustment on one GPU (the current CUDA device at construction)."""

    def __init__(self, prob: Problem, plan: int = 0, group=None):
        """``plan``: feba_settings.plan.  ``group`` = (rank, world, unique_id): one rank of a group of GPUs working
        on ONE adjustment (feba_create_shard; every rank passes the complete problem)."""
        self._lib = load()
        self._h = C.c_void_p()
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        # keep the arrays alive for the duration of the create call
        keep = dict(x=f64(prob.obs_x), y=f64(prob.obs_y), im=i32(prob.obs_img), pt=i32(prob.obs_pt),
                    ic=i32(prob.img_cam), eop=f64(prob.eop0), iop=f64(prob.iop0), box=f64(prob.cam_box),
                    xyz=f64(prob.xyz0), tie=i32(prob.pt_tie))
        fp = FebaProblem()
        fp.n_obs, fp.n_img, fp.n_cam = prob.n_obs, prob.numImg, prob.numCam
        fp.n_pts, fp.n_tie = prob.numPts, prob.numtie
        fp.obs_x, fp.obs_y, fp.obs_img, fp.obs_pt = _dp(keep["x"]), _dp(keep["y"]), _ip(keep["im"]), _ip(keep["pt"])
        fp.img_cam, fp.eop0, fp.iop0, fp.cam_box = _ip(keep["ic"]), _dp(keep["eop"]), _dp(keep["iop"]), _dp(keep["box"])
        fp.xyz0, fp.pt_tie = _dp(keep["xyz"]), _ip(keep["tie"])
        fp.settings = settings_struct(prob.settings, plan)
        if keep["iop"].shape != (prob.numCam, 3 + prob.settings.NK + 2):
            raise FebaError(FEBA_ERR_INVALID, "iop0 must be numCam x (3+NK+2)")
        if group is not None and group[1] > 1:
            rank, world, uid = group
            buf = C.create_string_buffer(bytes(uid), DIST_ID_BYTES)
            rc = self._lib.feba_create_shard(C.byref(fp), rank, world, buf, DIST_ID_BYTES, C.byref(self._h))
        else:
            rc = self._lib.feba_create(C.byref(fp), C.byref(self._h))
        if rc != 0:
            text = self._lib.feba_last_error(None).decode()
            self._h = C.c_void_p()
            raise FebaError(rc, text)
        u, uc = C.c_int64(), C.c_int64()
        self._lib.feba_num_unknowns(self._h, C.byref(u), C.byref(uc))
        self.u, self.u_c = int(u.value), int(uc.value)
        self.n_obs = int(self._lib.feba_num_obs(self._h))      # rows of the PHO table (global for a group)
        assert self.n_obs == prob.n_obs

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.feba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise FebaError(rc, self._lib.feba_last_error(self._h).decode())

    # -- calls
    def set_stream(self, cuda_stream: int):
        self._check(self._lib.feba_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_xhat(self, xhat: np.ndarray):
        x = np.ascontiguousarray(xhat, dtype=np.float64)
        self._check(self._lib.feba_set_xhat(self._h, _dp(x), x.size))

    def get_xhat(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        x = np.empty(self.u, dtype=np.float64) if out is None else out
        self._check(self._lib.feba_get_xhat(self._h, _dp(x), x.size))
        return x

    def set_xhat_owned(self, xhat: np.ndarray):
        """Group handle: upload the EOP/IOP part and the tie points this rank owns only (feba_set_xhat_owned)."""
        x = np.ascontiguousarray(xhat, dtype=np.float64)
        self._check(self._lib.feba_set_xhat_owned(self._h, _dp(x), x.size))

    def get_xhat_owned(self, out: np.ndarray) -> np.ndarray:
        """Group handle: write the EOP/IOP part and the tie points this rank owns into ``out`` (not collective)."""
        self._check(self._lib.feba_get_xhat_owned(self._h, _dp(out), out.size))
        return out

    def num_owned_ties(self) -> int:
        return int(self._lib.feba_num_owned_ties(self._h))

    def get_delta(self) -> np.ndarray:
        d = np.empty(self.u, dtype=np.float64)
        self._check(self._lib.feba_get_delta(self._h, _dp(d), d.size))
        return d

    def iterate(self) -> float:
        ds = C.c_double()
        self._check(self._lib.feba_iterate(self._h, C.byref(ds)))
        return float(ds.value)

    def iterate_async(self):
        self._check(self._lib.feba_iterate_async(self._h))

    def iterate_solve_async(self):
        self._check(self._lib.feba_iterate_solve_async(self._h))

    def sync(self) -> float:
        ds = C.c_double()
        self._check(self._lib.feba_sync(self._h, C.byref(ds)))
        return float(ds.value)

    def iterate_assemble(self):
        self._check(self._lib.feba_iterate_assemble(self._h))

    def reduced_dev(self):
        """(device pointer, number of doubles) of this rank's reduced system buffer."""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.feba_reduced_dev(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def reduced_pack(self):
        """feba_reduced_pack: copy what the solve half reads of the reduced system into one contiguous
        device buffer; returns (device pointer, number of doubles).  Pair with ``reduced_unpack``."""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.feba_reduced_pack(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def reduced_unpack(self):
        self._check(self._lib.feba_reduced_unpack(self._h))

    def dist_init(self, rank: int, world: int, unique_id: bytes):
        """Join a group of handles (one per GPU/process) that factorise the reduced system together
        (feba_dist_init; collective).  ``unique_id`` comes from ``dist_unique_id()`` on one rank."""
        buf = C.create_string_buffer(bytes(unique_id), DIST_ID_BYTES)
        self._check(self._lib.feba_dist_init(self._h, rank, world, buf, DIST_ID_BYTES))

    def iterate_solve(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._lib.feba_iterate_solve(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def debug_reduced(self):
        S = np.empty((self.u_c, self.u_c), dtype=np.float64)
        g = np.empty(self.u_c, dtype=np.float64)
        self._check(self._lib.feba_debug_reduced(self._h, _dp(S), _dp(g)))
        return S, g

    def solve(self):
        """Whole loop main.m:412-494.  Returns (iterations, deltasum trace)."""
        cap = 4096
        it = C.c_int32()
        trace = np.zeros(cap, dtype=np.float64)
        self._check(self._lib.feba_solve(self._h, C.byref(it), _dp(trace), cap))
        return int(it.value), trace[:min(int(it.value), cap)].copy()

    def residuals(self, want_v: bool = True, want_rsd: bool = True, v_out: Optional[np.ndarray] = None,
                  rsd_out: Optional[np.ndarray] = None):
        """main.m:569-601 + BuildRSD.  Returns dict(v, RSD, RMSx, RMSy, RMS, sigma02, sxx, syy).
        ``v_out`` (2 n_obs) / ``rsd_out`` (n_obs x 5): caller-owned output arrays; page-locked ones (e.g. views of a
        pinned torch tensor) are written by one DMA instead of through the library's staging buffers."""
        v = (np.empty(2 * self.n_obs, dtype=np.float64) if v_out is None else v_out) if want_v else None
        rsd = (np.empty((self.n_obs, 5), dtype=np.float64) if rsd_out is None else rsd_out) if want_rsd else None
        st = np.zeros(6, dtype=np.float64)
        self._check(self._lib.feba_residuals(self._h, _dp(v) if want_v else None,
                                             _dp(rsd) if want_rsd else None, _dp(st)))
        return dict(v=v, RSD=rsd, RMSx=st[0], RMSy=st[1], RMS=st[2], sigma02=st[3], sxx=st[4], syy=st[5])

    def cov_diag(self) -> np.ndarray:
        """diag(Cx)/sigma02 of all unknowns (main.m:432-444, un-scaled as main.m:468-480)."""
        q = np.empty(self.u, dtype=np.float64)
        self._check(self._lib.feba_cov_diag(self._h, _dp(q), q.size))
        return q

    def cov_block(self, idx) -> np.ndarray:
        """k x k block of Cx/sigma02 before un-scaling for EOP/IOP unknown indices ``idx``."""
        ii = np.ascontiguousarray(idx, dtype=np.int64)
        out = np.empty((ii.size, ii.size), dtype=np.float64)
        self._check(self._lib.feba_cov_block(self._h, ii.ctypes.data_as(C.POINTER(C.c_int64)), ii.size, _dp(out)))
        return out

    def last_timing(self):
        ms = np.zeros(6, dtype=np.float64)
        self._check(self._lib.feba_last_timing(self._h, _dp(ms)))
        return dict(prep_ms=ms[0], assemble_ms=ms[1], factor_ms=ms[2], solve_ms=ms[3], update_ms=ms[4],
                    total_ms=ms[5])

    def last_timing_ex(self):
        ms = np.zeros(8, dtype=np.float64)
        self._check(self._lib.feba_last_timing_ex(self._h, _dp(ms)))
        return dict(prep_ms=ms[0], assemble_ms=ms[1], factor_ms=ms[2], solve_ms=ms[3], update_ms=ms[4],
                    total_ms=ms[5], exchange_ms=ms[6], residual_kernel_ms=ms[7])

    def plan_info(self) -> dict:
        """Row order / supertiles of the reduced system in use (feba_plan_info)."""
        v = (C.c_int32 * 8)()
        fl = np.zeros(2, dtype=np.float64)
        self._check(self._lib.feba_plan_info(self._h, v, _dp(fl)))
        return dict(nested_dissection=bool(v[0]), rows=int(v[1]), supertiles=int(v[2]), nodes=int(v[3]),
                    chain_blocks=int(v[4]), world=int(v[5]), top_row0=int(v[6]), local_obs=int(v[7]),
                    flop=float(fl[0]), flop_dense=float(fl[1]))

    def launch_count(self) -> int:
        return int(self._lib.feba_launch_count(self._h))

    def sparse_info(self) -> dict:
        """Block-sparse form of the reduced system (FEBA_SPARSE=1 at creation): active, non-zero / all lower
        supertiles (fill included), datum images."""
        v = (C.c_int32 * 4)()
        self._check(self._lib.feba_sparse_info(self._h, v))
        return dict(active=bool(v[0]), nonzero_supertiles=int(v[1]), lower_supertiles=int(v[2]), datum_images=int(v[3]))


class Batch:
    """feba_batch: independent handles on one device advancing together, one CUDA graph launch per step."""

    def __init__(self, handles):
        self._lib = load()
        self.handles = list(handles)
        arr = (C.c_void_p * len(self.handles))(*[h._h for h in self.handles])
        self._b = C.c_void_p()
        rc = self._lib.feba_batch_create(arr, len(self.handles), C.byref(self._b))
        if rc != 0:
            raise FebaError(rc, self._lib.feba_last_error(None).decode())

    def iterate_async(self):
        rc = self._lib.feba_batch_iterate_async(self._b)
        if rc != 0:
            raise FebaError(rc, self._lib.feba_last_error(self.handles[0]._h).decode())

    def close(self):
        if getattr(self, "_b", None) is not None and self._b.value:
            self._lib.feba_batch_destroy(self._b)
            self._b = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
