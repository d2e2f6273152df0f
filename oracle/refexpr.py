"""ORACLE TOOLING (test infrastructure) -- run the reference's own BuildAwG statements.

MATLAB is absent from the build container, so the reference cannot be executed as a
program.  Its per-observation arithmetic, however, is plain scalar MATLAB whose only
non-Python tokens are ``^``, 1-based ``K(j)`` / ``P(1)`` indexing and ``atan``/``sec``.  This
module reads ``functions/BuildAwG.m`` *from the reference tree at run time* (nothing is
copied into this repo), harvests the statements by the name on their left-hand side and the
model branch they sit in, rewrites those tokens, and evaluates them with NumPy over all
observations at once.  It yields what one pass of the reference loop body
(``BuildAwG.m:46-528``) would put into ``A``, ``misclosure``, ``G`` and ``dist_scaling``.

Used by ``tests/test_oracle_vs_reference_source.py`` (skipped when the reference tree is not
mounted, e.g. on the GPU box) and by ``tests/golden/make_golden.py`` which freezes the outputs
on the bundled cam0 data into ``tests/golden/*.npz``.
"""
from __future__ import annotations

import os
import re

import numpy as np

REFERENCE_ROOT = os.environ.get("FEBA_REFERENCE_ROOT", "/root/reference")
TYPE_NAMES = ("fisheye", "pinhole", "equisolid", "orthographic", "stereographic")

_GEN = r"A1[1-6]|A2[1-6]|Ax_c|Ay_c|dx_dX|dx_dY|dx_dZ|dy_dX|dy_dY|dy_dZ"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "functions", "BuildAwG.m"))


def _strip(line: str) -> str:
    p = line.find("%")
    if p >= 0:
        line = line[:p]
    return line.strip()


def _py(expr: str) -> str:
    """MATLAB scalar arithmetic -> Python/NumPy arithmetic."""
    e = expr.strip().rstrip(";").strip()
    e = e.replace("./", "/").replace(".*", "*").replace("^", "**")
    e = re.sub(r"data\.points\(i\)\.(\w+)", r"\1", e)
    e = re.sub(r"data\.settings\.Num_Radial_Distortions", "NK", e)
    e = re.sub(r"dist_scaling\(cam_num,([^)]+)\)", r"dist_scaling[(\1)-1]", e)
    e = re.sub(r"\b(K|P)\(([^()]+)\)", r"\1[(\2)-1]", e)
    return e


class ReferenceBuildAwG:
    """Harvested statements of BuildAwG.m, ready to evaluate."""

    def __init__(self, path: str | None = None):
        path = path or os.path.join(REFERENCE_ROOT, "functions", "BuildAwG.m")
        with open(path, "r") as fh:
            raw = fh.read().splitlines()
        self.lines = [_strip(l) for l in raw]
        self.gen = {}            # (name, typeint) -> python expr
        self.proj = {}           # (fx|fy, typeint) -> python expr
        self.simple = {}         # name -> python expr (first occurrence)
        self.loops = {}          # tag -> list of python statements executed for j=1..NK
        self.iop_rows = {}       # 'xp'|'yp' -> [row_x_expr, row_y_expr]
        self.grows = []          # 6 rows x 7 exprs
        self._harvest()

    def _harvest(self):
        typeint, model, section = None, None, None
        L = self.lines
        i = 0
        while i < len(L):
            s = L[i]
            m = re.match(r"(?:if|elseif)\s+typeint\s*==\s*(\d)", s)
            if m:
                typeint = int(m.group(1))
            m = re.match(r"(?:if|elseif)\s+strcmp\(data\.settings\.type,'(\w+)'\)", s)
            if m:
                model = TYPE_NAMES.index(m.group(1))
            m = re.match(r"if data\.settings\.Estimate_(xp|yp)\b", s)
            if m:
                section = m.group(1)
            m = re.match(rf"({_GEN})\s*=\s*(.+)$", s)
            if m and typeint is not None:
                self.gen[(m.group(1), typeint)] = _py(m.group(2))
            m = re.match(r"(fx|fy)\s*=\s*(.+)$", s)
            if m and model is not None:
                self.proj[(m.group(1), model)] = _py(m.group(2))
            m = re.match(r"(U|V|W|R|x_bar|y_bar|r|decentering_x|decentering_y|rmax|w11|w21|Ax_P|Ay_P)"
                         r"\s*=\s*(.+)$", s)
            if m and m.group(1) not in self.simple:
                rhs = m.group(2)
                if m.group(1) in ("Ax_P", "Ay_P"):
                    # "[a b]./scale" -> two expressions
                    mm = re.match(r"\[(.+)\]\./(.+?);?$", rhs)
                    parts = self._split_row(mm.group(1))
                    self.simple[m.group(1)] = [_py(f"({p})/({mm.group(2)})") for p in parts]
                else:
                    self.simple[m.group(1)] = _py(rhs)
            m = re.match(r"delta_r\s*=\s*delta_r\s*\+(.+)$", s)
            if m:
                self.loops["delta_r"] = _py("delta_r +" + m.group(1))
            m = re.match(r"(dxp_rad|dyp_rad)\s*=\s*\1(.+)$", s)
            if m and section:
                self.loops[(section, m.group(1))] = _py(m.group(1) + m.group(2))
            m = re.match(r"(Ax_K|Ay_K)\(1,j\)\s*=\s*(.+)$", s)
            if m:
                self.loops[m.group(1)] = _py(m.group(2))
            m = re.match(r"dist_scaling\(data\.points\(i\)\.cam_num,2\+j\)\s*=\s*(.+)$", s)
            if m:
                self.loops["scale"] = _py(m.group(1))
            if s.startswith("Ablock_IOPs(:,count_A) = [") and section and "Ax_c" not in s:
                row_x = s[s.index("[") + 1:].rstrip(";").strip()
                row_y = L[i + 1].replace("];", "").rstrip(";").strip()
                self.iop_rows[section] = [_py(row_x), _py(row_y)]
            if s.startswith("Gblock = ["):
                for r in range(6):
                    toks = L[i + 1 + r].rstrip(";").split()
                    assert len(toks) == 7, toks
                    self.grows.append([_py(t) for t in toks])
            i += 1
        # dist_scaling(data.points(i).cam_num, ...) inside expressions
        fix = lambda e: re.sub(r"dist_scaling\(cam_num,([^)]+)\)", r"dist_scaling[(\1)-1]", e)
        for k in list(self.loops):
            self.loops[k] = fix(self.loops[k])

    @staticmethod
    def _split_row(text: str):
        """Split a MATLAB row literal on top-level blanks: '(a + b) 2*x*y' -> 2 items."""
        out, depth, cur = [], 0, ""
        for ch in text.strip():
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == " " and depth == 0:
                if cur:
                    out.append(cur)
                cur = ""
            else:
                cur += ch
        if cur:
            out.append(cur)
        return out

    # ------------------------------------------------------------------ evaluation
    def evaluate(self, typeint, NK, x, y, Xc, Yc, Zc, w, p, k, X, Y, Z, xp, yp, c, K, P, y_dir,
                 xmin, ymin, xmax, ymax):
        """One pass of the loop body for arrays of observations.  ``K``: list of NK arrays,
        ``P``: list of 2 arrays.  Returns fx, fy, w, Je (n,2,6), Jc (n,2,3+NK+2), Jt (n,2,3),
        scale (n,NK)."""
        env = dict(sin=np.sin, cos=np.cos, tan=np.tan, atan=np.arctan, sqrt=np.sqrt,
                   sec=lambda a: 1.0 / np.cos(a), x=x, y=y, Xc=Xc, Yc=Yc, Zc=Zc, w=w, p=p, k=k,
                   X=X, Y=Y, Z=Z, xp=xp, yp=yp, c=c, K=K, P=P, y_dir=y_dir, xmin=xmin,
                   ymin=ymin, xmax=xmax, ymax=ymax, NK=NK)
        ev = lambda e: eval(e, {"__builtins__": {}}, env)
        for nm in ("U", "V", "W", "R", "x_bar", "y_bar", "r"):
            env[nm] = ev(self.simple[nm])
        env["delta_r"] = np.zeros_like(x)
        for j in range(1, NK + 1):
            env["j"] = j
            env["delta_r"] = ev(self.loops["delta_r"])
        env["decentering_x"] = ev(self.simple["decentering_x"])
        env["decentering_y"] = ev(self.simple["decentering_y"])
        fx = ev(self.proj[("fx", typeint)])
        fy = ev(self.proj[("fy", typeint)])
        env["fx"], env["fy"] = fx, fy
        n = x.shape[0]
        Je = np.empty((n, 2, 6))
        # column order Xc Yc Zc w p k  <->  A14/A24, A15/A25, A16/A26, A11/A21, A12/A22, A13/A23
        for col, tag in enumerate(("4", "5", "6", "1", "2", "3")):
            Je[:, 0, col] = ev(self.gen[("A1" + tag, typeint)])
            Je[:, 1, col] = ev(self.gen[("A2" + tag, typeint)])
        Jt = np.empty((n, 2, 3))
        for col, ax in enumerate("XYZ"):
            Jt[:, 0, col] = ev(self.gen[("dx_d" + ax, typeint)])
            Jt[:, 1, col] = ev(self.gen[("dy_d" + ax, typeint)])
        Jc = np.zeros((n, 2, 3 + NK + 2))
        for col, sec_ in enumerate(("xp", "yp")):
            env["dxp_rad"] = np.zeros_like(x)
            env["dyp_rad"] = np.zeros_like(x)
            for j in range(1, NK + 1):
                env["j"] = j
                # MATLAB executes the two statements in sequence (they are independent)
                a = ev(self.loops[(sec_, "dxp_rad")])
                b = ev(self.loops[(sec_, "dyp_rad")])
                env["dxp_rad"], env["dyp_rad"] = a, b
            Jc[:, 0, col] = ev(self.iop_rows[sec_][0])
            Jc[:, 1, col] = ev(self.iop_rows[sec_][1])
        Jc[:, 0, 2] = ev(self.gen[("Ax_c", typeint)])
        Jc[:, 1, 2] = ev(self.gen[("Ay_c", typeint)])
        env["rmax"] = ev(self.simple["rmax"])
        scale = np.empty((n, NK))
        for j in range(1, NK + 1):
            env["j"] = j
            scale[:, j - 1] = ev(self.loops["scale"])
        env["dist_scaling"] = [None, None] + [scale[:, j] for j in range(NK)]
        for j in range(1, NK + 1):
            env["j"] = j
            Jc[:, 0, 2 + j] = ev(self.loops["Ax_K"])
            Jc[:, 1, 2 + j] = ev(self.loops["Ay_K"])
        for t in range(2):
            Jc[:, 0, 3 + NK + t] = ev(self.simple["Ax_P"][t])
            Jc[:, 1, 3 + NK + t] = ev(self.simple["Ay_P"][t])
        wv = np.stack([ev(self.simple["w11"]), ev(self.simple["w21"])], axis=-1)
        return dict(fx=fx, fy=fy, w=wv, Je=Je, Jc=Jc, Jt=Jt, scale=scale)

    def evaluate_G(self, Xc, Yc, Zc, w, p):
        env = dict(sin=np.sin, cos=np.cos, tan=np.tan, sec=lambda a: 1.0 / np.cos(a),
                   Xc=Xc, Yc=Yc, Zc=Zc, w=w, p=p)
        n = np.shape(Xc)[0]
        G = np.zeros((n, 6, 7))
        for r in range(6):
            for cidx in range(7):
                G[:, r, cidx] = eval(self.grows[r][cidx], {"__builtins__": {}}, env)
        return G


def reference_observation_equations(prob, eop, iop, xyz, ref: ReferenceBuildAwG | None = None):
    """Same outputs as ``oracle.model.observation_equations`` but computed by the reference's
    own statements."""
    ref = ref or ReferenceBuildAwG()
    s = prob.settings
    NK = s.NK
    im, pt = prob.obs_img, prob.obs_pt
    cam = prob.img_cam[im]
    box = prob.cam_box[cam]
    out = ref.evaluate(
        s.typeint, NK, prob.obs_x, prob.obs_y,
        *(eop[im, q] for q in range(6)), xyz[pt, 0], xyz[pt, 1], xyz[pt, 2],
        iop[cam, 0], iop[cam, 1], iop[cam, 2], [iop[cam, 3 + j] for j in range(NK)],
        [iop[cam, 3 + NK], iop[cam, 4 + NK]], box[:, 0], box[:, 1], box[:, 2], box[:, 3], box[:, 4])
    out["G"] = ref.evaluate_G(*(eop[:, q] for q in (0, 1, 2, 3, 4)))
    return out
