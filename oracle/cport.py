"""ORACLE (test infrastructure, not product code) -- Python front of the C / OpenMP restatement
(``oracle/feba_oracle.c``): same Gauss-Newton step as ``oracle/sparse.py`` but fast enough for the
BASELINE.json sizes.  The dense reduced solve goes through LAPACK (``scipy.linalg``), the rest runs in
C.  Used as (a) a second, independently written checker for the CUDA path at sizes the NumPy oracle
cannot hold, and (b) the timed CPU baseline of ``bench.py`` ("port", all host cores).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np
import scipy.linalg as sla

from .model import G_rows, gather_params, layout

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfeba_oracle.so")
_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int)


class OracleProblem(C.Structure):
    _fields_ = [("n_obs", C.c_long), ("n_img", C.c_int), ("n_cam", C.c_int), ("n_pts", C.c_int), ("n_tie", C.c_int),
                ("ox", _pd), ("oy", _pd), ("oimg", _pi), ("opt", _pi), ("img_cam", _pi), ("pt_tie", _pi),
                ("eop", _pd), ("iop", _pd), ("cam_box", _pd), ("xyz", _pd), ("ecol", _pi), ("ccol", _pi),
                ("NK", C.c_int), ("type", C.c_int), ("ui", C.c_int), ("uc", C.c_int), ("px", C.c_double),
                ("py", C.c_double), ("pt_start", _pi), ("pt_obs", _pi)]


def _host_signature() -> str:
    """CPU model + ISA flags of this machine: the library is compiled with -march=native, so a copy built on
    another machine (the repository snapshot travels between boxes) must be rebuilt."""
    try:
        with open("/proc/cpuinfo") as fh:
            txt = fh.read()
        model = next((ln for ln in txt.splitlines() if ln.startswith("model name")), "")
        flags = next((ln for ln in txt.splitlines() if ln.startswith("flags")), "")
        import hashlib
        return hashlib.sha1((model + flags).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "feba_oracle.c")
    mk = os.path.join(HERE, "Makefile")
    sig_file = os.path.join(HERE, "_build", "host.sig")
    sig = _host_signature()
    try:
        same_host = open(sig_file).read().strip() == sig
    except OSError:
        same_host = False
    stale = (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(mk)))
    if force or stale or not same_host:
        subprocess.run(["make", "-B", "-C", HERE], check=True, stdout=subprocess.DEVNULL)
        with open(sig_file, "w") as fh:
            fh.write(sig)
    return LIB


def set_threads(n: int) -> int:
    """OpenMP threads of the C restatement AND the BLAS/LAPACK threads of the dense solve; returns the OpenMP count
    in effect (bench.py: all host cores, also under a launcher that exported OMP_NUM_THREADS=1)."""
    lib().feba_oracle_set_threads(int(n))
    try:
        from threadpoolctl import threadpool_limits
        set_threads._blas = threadpool_limits(limits=int(n), user_api="blas")
    except Exception:        # threadpoolctl missing: BLAS keeps its own default
        pass
    return int(lib().feba_oracle_threads())


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.feba_oracle_assemble.argtypes = [C.POINTER(OracleProblem), _pd, C.c_long, _pd, _pd, _pd]
        _lib.feba_oracle_backsub.argtypes = [C.POINTER(OracleProblem), _pd, _pd, _pd, _pd]
        _lib.feba_oracle_residuals.argtypes = [C.POINTER(OracleProblem), _pd, _pd, _pd]
        _lib.feba_oracle_obs.argtypes = [C.c_int, C.c_int, C.c_long] + [_pd] * 10
        _lib.feba_oracle_set_threads.argtypes = [C.c_int]
    return _lib


def _dp(a):
    return a.ctypes.data_as(_pd)


def _ip(a):
    return a.ctypes.data_as(_pi)


class CPort:
    """Holds the static index arrays of one problem for the C calls."""

    def __init__(self, prob):
        self.prob = prob
        self.L = layout(prob)
        s = prob.settings
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self.ox, self.oy = f64(prob.obs_x), f64(prob.obs_y)
        self.oimg, self.opt = i32(prob.obs_img), i32(prob.obs_pt)
        self.img_cam, self.pt_tie = i32(prob.img_cam), i32(prob.pt_tie)
        self.cam_box = f64(prob.cam_box)
        self.ecol, self.ccol = i32(self.L["ecols"]), i32(self.L["ccols"])
        order = np.argsort(self.opt, kind="stable").astype(np.int32)
        self.pt_obs = order
        cnt = np.bincount(self.opt, minlength=prob.numPts)
        self.pt_start = i32(np.concatenate([[0], np.cumsum(cnt)]))
        self.px, self.py = 1.0 / s.sigma_x ** 2, 1.0 / s.sigma_y ** 2
        box = prob.cam_box
        self.rmax2 = ((box[:, 3] - box[:, 1]) * 0.5) ** 2 + ((box[:, 4] - box[:, 2]) * 0.5) ** 2

    def _struct(self, eop, iop, xyz):
        p, s, L = self.prob, self.prob.settings, self.L
        st = OracleProblem()
        st.n_obs, st.n_img, st.n_cam, st.n_pts, st.n_tie = p.n_obs, p.numImg, p.numCam, p.numPts, p.numtie
        st.ox, st.oy, st.oimg, st.opt = _dp(self.ox), _dp(self.oy), _ip(self.oimg), _ip(self.opt)
        st.img_cam, st.pt_tie = _ip(self.img_cam), _ip(self.pt_tie)
        st.eop, st.iop, st.cam_box, st.xyz = _dp(eop), _dp(iop), _dp(self.cam_box), _dp(xyz)
        st.ecol, st.ccol = _ip(self.ecol), _ip(self.ccol)
        st.NK, st.type, st.ui, st.uc = s.NK, s.typeint, L["u_img"], L["u_cam"]
        st.px, st.py = self.px, self.py
        st.pt_start, st.pt_obs = _ip(self.pt_start), _ip(self.pt_obs)
        return st

    def iterate(self, xhat, timing=None, diag_shift=0.0):
        """One Gauss-Newton step (main.m:416-488).  Returns xhat_new, deltasum, state."""
        p, s, L = self.prob, self.prob.settings, self.L
        u_c = L["off_tie"]
        t0 = time.perf_counter()
        eop, iop, xyz = (np.ascontiguousarray(a) for a in gather_params(p, xhat))
        st = self._struct(eop, iop, xyz)
        S = np.zeros((u_c, u_c), order="F")
        g = np.zeros(u_c)
        Vinv = np.zeros((max(p.numtie, 1), 6))
        up = np.zeros((max(p.numtie, 1), 3))
        rc = lib().feba_oracle_assemble(C.byref(st), _dp(S), u_c, _dp(g), _dp(Vinv), _dp(up))
        if rc:
            raise RuntimeError("oracle: singular point block")
        t1 = time.perf_counter()
        if diag_shift:                      # timing runs on thinned samples only: keep S safely definite
            S[np.diag_indices_from(S)] += diag_shift * np.abs(np.diag(S)).max()
        # (bordered) dense solve on the lower triangle: LAPACK dpotrf / dpotrs
        if s.Inner_Constraints:
            Gi = G_rows(eop)
            Gc = np.zeros((u_c, 7))
            for j in np.unique(p.obs_img):
                Gc[6 * j:6 * j + 6] = Gi[j]
            M = S
            M += np.tril(Gc @ Gc.T)
            cf = sla.cho_factor(M, lower=True, overwrite_a=True, check_finite=False)
            Y = sla.cho_solve(cf, np.column_stack([g, Gc]), check_finite=False)
            k = np.linalg.solve(Gc.T @ Y[:, 1:], -(Gc.T @ Y[:, 0]))
            d_c = -(Y[:, 0] + Y[:, 1:] @ k)
        else:
            cf = sla.cho_factor(S, lower=True, overwrite_a=True, check_finite=False)
            d_c = -sla.cho_solve(cf, g, check_finite=False)
        t2 = time.perf_counter()
        dpts = np.zeros((max(p.numtie, 1), 3))
        if p.numtie:
            lib().feba_oracle_backsub(C.byref(st), _dp(np.ascontiguousarray(d_c)), _dp(Vinv), _dp(up), _dp(dpts))
        delta_s = np.concatenate([d_c, dpts[:p.numtie].reshape(-1)])
        delta = delta_s.copy()                                  # un-scaling, main.m:458-482
        for c in range(p.numCam):
            base = L["off_cam"] + L["u_cam"] * c
            if s.Estimate_radial:
                for j in range(s.Num_Radial_Distortions):
                    delta[base + L["ccols"][3] + j] /= self.rmax2[c] ** (j + 1)
            if s.Estimate_decent:
                for j in range(2):
                    delta[base + L["ccols"][3 + L["NK"]] + j] /= self.rmax2[c]
        t3 = time.perf_counter()
        if timing is not None:
            timing.update(assemble_s=t1 - t0, solve_s=t2 - t1, backsub_s=t3 - t2, total_s=t3 - t0)
        deltasum = float(np.sum(np.abs(delta)))
        state = dict(eop=eop, iop=iop, xyz=xyz, delta=delta, dpts=dpts)
        return xhat + delta, deltasum, state

    def residuals(self, state, xhat_new):
        """main.m:569 + BuildRSD + main.m:592-602 with the tables of the LAST linearisation point."""
        from .dense import BuildRSD
        p, L = self.prob, self.L
        st = self._struct(state["eop"], state["iop"], state["xyz"])
        v = np.zeros(2 * p.n_obs)
        dcu = np.ascontiguousarray(state["delta"][:L["off_tie"]])
        lib().feba_oracle_residuals(C.byref(st), _dp(dcu), _dp(state["dpts"]), _dp(v))
        RSD = BuildRSD(p, v, xhat_new)
        vx, vy = v[0::2], v[1::2]
        RMSx, RMSy = np.sqrt(np.mean(vx ** 2)), np.sqrt(np.mean(vy ** 2))
        sigma02 = float(np.sum(vx ** 2) * self.px + np.sum(vy ** 2) * self.py) / (p.n - L["u"])
        return dict(v=v, RSD=RSD, RMSx=RMSx, RMSy=RMSy, RMS=np.sqrt(RMSx ** 2 + RMSy ** 2), sigma02=sigma02)

    def gauss_newton(self, xhat0, max_iter=None):
        s = self.prob.settings
        xhat = np.array(xhat0, dtype=np.float64).copy()
        deltasum, count, trace, st = 100.0, 0, [], None
        cap = s.Iteration_Cap if max_iter is None else max_iter
        while deltasum > s.threshold:
            count += 1
            xhat, deltasum, st = self.iterate(xhat)
            trace.append(deltasum)
            if count >= cap:
                break
        out = self.residuals(st, xhat)
        out.update(xhat=xhat, iterations=count, deltasum=trace, delta=st["delta"])
        return out


def observation_equations(prob, eop, iop, xyz):
    """Per-observation Jacobians from the C restatement (same outputs as oracle.model)."""
    s = prob.settings
    NK, NC = s.NK, s.NK + 5
    im, pt = prob.obs_img, prob.obs_pt
    cam = prob.img_cam[im]
    n = prob.n_obs
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    e, i_, b, x = f64(eop[im]), f64(iop[cam]), f64(prob.cam_box[cam]), f64(xyz[pt])
    Je, Jc, Jt, w = np.zeros((n, 2, 6)), np.zeros((n, 2, NC)), np.zeros((n, 2, 3)), np.zeros((n, 2))
    lib().feba_oracle_obs(s.typeint, NK, n, _dp(f64(prob.obs_x)), _dp(f64(prob.obs_y)), _dp(e), _dp(i_), _dp(b),
                          _dp(x), _dp(Je), _dp(Jc), _dp(Jt), _dp(w))
    return dict(Je=Je, Jc=Jc, Jt=Jt, w=w)
