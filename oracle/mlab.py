"""ORACLE TOOLING (test infrastructure, never on the product path) -- execute the reference's own
MATLAB source with a small MATLAB-subset interpreter.

MATLAB / Octave are absent from the build image, so the reference cannot run as a program.  Its
numerical core, however, is written in a narrow, plain subset of the language: scalar and matrix
arithmetic, ``for`` / ``while`` / ``if``, 1-based indexing with ranges, matrix literals, struct fields,
a few cells and a dozen built-ins.  This module transpiles that subset to Python at run time and
executes it on NumPy, so that whole FUNCTIONS of the reference -- ``functions/BuildAwG.m``,
``functions/Buildxhat.m``, ``functions/BuildRSD.m``, ``functions/sumabs.m`` -- and the Gauss-Newton
loop, residual and statistics statements of ``main.m`` (``:396-494``, ``:569``, ``:592-602`` and the
local ``rms``) run *as written*, from the files under ``/root/reference`` (nothing is copied into this
repository; without the tree the users of this module skip).

``oracle/refexpr.py`` checks the oracle expression by expression; this one checks it function by
function and loop by loop: index arithmetic, block placement in ``A``, ``dist_scaling``, ``G`` rows,
the bordered inverse, un-scaling, ``sumabs``, ``v = A*delta + w``, ``BuildRSD``, RMS, ``sigma02``.
Frozen outputs on the bundled data: ``tests/golden/cam0_refrun_*.npz`` (``tests/golden/make_golden.py``).

What it is not: MATLAB.  Elementary functions come from libm, ``^-1`` is ``numpy.linalg.inv`` (LAPACK
``dgetrf/dgetri`` like MATLAB's ``inv``, but not the same build), ``x^2`` is ``pow``.  Differences of a
few ulp per entry are expected and are amplified by cond(N) in the solve; the tests state the
tolerances they use.

Semantics implemented (everything else raises ``MlabError`` rather than guessing):
values are Python floats (scalars), ``Mat`` (2-D double), ``str``, ``Cell``, ``Struct`` / ``StructArray``;
``*`` is the matrix product unless one side is scalar, ``/`` needs a scalar divisor, ``^`` is scalar
power or ``^-1`` = inverse; ``'`` transposes; indexing is 1-based, column-major for one subscript,
with ``:``, ``a:b``, ``a:s:b`` and ``end``; matrix literals split elements on blanks the way MATLAB
does (``[1 -x]`` is two elements, ``[1 - x]`` one); statements that only talk to the user
(``disp``, ``errordlg``, ``waitfor``, ``tic``) are dropped; ``fprintf`` writes into a ``FileSink``
(escapes, positional ``%N$`` conversions, cyclic reuse of the format).
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict, List, Optional, Sequence

import numpy as np

REFERENCE_ROOT = os.environ.get("FEBA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "functions", "BuildAwG.m"))


class MlabError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------ values

class _Colon:
    pass


COLON = _Colon()


class _End:
    """``end`` inside a subscript, with ``end-1`` / ``end+k`` arithmetic."""

    def __init__(self, off: int = 0):
        self.off = off

    def __sub__(self, k):
        return _End(self.off - int(round(float(k))))

    def __add__(self, k):
        return _End(self.off + int(round(float(k))))

    __radd__ = __add__

    def at(self, n: int) -> int:
        return n + self.off


END = _End()


class Rng:
    """``a:b`` / ``a:s:b`` (``b`` may be ``end``)."""

    def __init__(self, a, b, c=None):
        self.a, self.s, self.b = (a, 1, b) if c is None else (a, b, c)

    def resolve(self, n: int) -> np.ndarray:
        b = self.b.at(n) if isinstance(self.b, _End) else self.b
        a = self.a.at(n) if isinstance(self.a, _End) else self.a
        return np.arange(int(round(a)), int(round(b)) + (1 if self.s > 0 else -1), int(round(self.s)))

    def __iter__(self):
        if isinstance(self.b, _End) or isinstance(self.a, _End):
            raise MlabError("'end' outside an index")
        return iter(int(v) for v in self.resolve(0))

    def mat(self) -> "Mat":
        return Mat(self.resolve(0).astype(float).reshape(1, -1))


def _num(v) -> float:
    if isinstance(v, Mat):
        if v.a.size != 1:
            raise MlabError("matrix used where a scalar is needed")
        return float(v.a.flat[0])
    return float(v)


def _index(i, n: int):
    """One subscript -> 0-based integer array or int."""
    if i is COLON:
        return np.arange(n)
    if isinstance(i, Rng):
        return i.resolve(n) - 1
    if isinstance(i, _End):
        return i.at(n) - 1
    if isinstance(i, Mat):
        return np.rint(i.a.ravel(order="F")).astype(int) - 1
    k = int(round(float(i)))
    if abs(k - float(i)) > 1e-9 or k < 1:
        raise MlabError(f"bad subscript {i!r}")
    return k - 1


class Mat:
    """2-D double matrix with MATLAB operator semantics."""

    __array_priority__ = 100
    __slots__ = ("a",)

    def __init__(self, a):
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(1, -1)
        self.a = a

    # ---- indexing
    def __call__(self, *idx):
        if len(idx) == 1:
            flat = self.a.ravel(order="F")
            k = _index(idx[0], flat.size)
            if isinstance(k, (int, np.integer)):
                return float(flat[k])
            out = flat[k]
            if self.a.shape[0] == 1 and self.a.shape[1] != 1:      # row vector stays a row
                return Mat(out.reshape(1, -1))
            if self.a.shape[1] == 1:                                # column vector stays a column
                return Mat(out.reshape(-1, 1))
            if isinstance(idx[0], Mat):                             # matrix source: shape of the subscript
                return Mat(out.reshape(idx[0].a.shape, order="F"))
            return Mat(out.reshape(1, -1) if isinstance(idx[0], Rng) else out.reshape(-1, 1))
        if len(idx) != 2:
            raise MlabError("only 1 or 2 subscripts")
        r, c = _index(idx[0], self.a.shape[0]), _index(idx[1], self.a.shape[1])
        if isinstance(r, (int, np.integer)) and isinstance(c, (int, np.integer)):
            return float(self.a[r, c])
        return Mat(self.a[np.ix_(np.atleast_1d(r), np.atleast_1d(c))])

    def set(self, idx, val):
        v = val.a if isinstance(val, Mat) else float(val)
        if len(idx) == 1:
            k = _index(idx[0], self.a.size)
            if np.max(k) >= self.a.size:
                raise MlabError("assignment would grow the array (not supported)")
            r, c = np.unravel_index(k, self.a.shape, order="F")
            self.a[r, c] = v.ravel(order="F") if isinstance(v, np.ndarray) and v.size > 1 else (
                float(v.flat[0]) if isinstance(v, np.ndarray) else v)
            return
        r, c = _index(idx[0], self.a.shape[0]), _index(idx[1], self.a.shape[1])
        if np.size(r) == 0 or np.size(c) == 0:                  # A(1:2, 5:4) = zeros(2,0): nothing to store
            if isinstance(v, np.ndarray) and v.size:
                raise MlabError("non-empty value into an empty selection")
            return
        if np.max(r) >= self.a.shape[0] or np.max(c) >= self.a.shape[1]:
            raise MlabError("assignment would grow the array (not supported)")
        if isinstance(r, (int, np.integer)) and isinstance(c, (int, np.integer)):
            self.a[r, c] = float(v.flat[0]) if isinstance(v, np.ndarray) else v
        else:
            rr, cc = np.atleast_1d(r), np.atleast_1d(c)
            if isinstance(v, np.ndarray) and v.size > 1 and v.shape != (rr.size, cc.size):
                raise MlabError(f"shape mismatch in assignment: {v.shape} into {(rr.size, cc.size)}")
            self.a[np.ix_(rr, cc)] = v

    @property
    def T(self):
        return Mat(self.a.T.copy())

    # ---- arithmetic
    @staticmethod
    def _s(x):
        """scalar value of x or None"""
        if isinstance(x, Mat):
            return float(x.a.flat[0]) if x.a.size == 1 else None
        return float(x)

    def _ew(self, other, op):
        o = other.a if isinstance(other, Mat) else float(other)
        if isinstance(o, np.ndarray) and o.size > 1 and self.a.size > 1 and o.shape != self.a.shape:
            raise MlabError(f"size mismatch {self.a.shape} vs {o.shape}")
        return Mat(op(self.a, o))

    def __add__(self, o): return self._ew(o, np.add)
    def __radd__(self, o): return self._ew(o, np.add)
    def __sub__(self, o): return self._ew(o, np.subtract)
    def __rsub__(self, o): return Mat(float(o) - self.a)
    def __neg__(self): return Mat(-self.a)
    def __pos__(self): return self

    def __mul__(self, o):
        if Mat._s(o) is not None:
            return Mat(self.a * Mat._s(o))
        if self.a.size == 1:
            return Mat(float(self.a.flat[0]) * o.a)
        if self.a.shape[1] != o.a.shape[0]:
            raise MlabError(f"inner dimensions {self.a.shape} * {o.a.shape}")
        return Mat(self.a @ o.a)

    def __rmul__(self, o): return Mat(float(o) * self.a)

    def __truediv__(self, o):
        s = Mat._s(o)
        if s is None:
            raise MlabError("matrix right-division is not in the subset")
        return Mat(self.a / s)

    def __rtruediv__(self, o):
        if self.a.size != 1:
            raise MlabError("division by a matrix is not in the subset")
        return float(o) / float(self.a.flat[0])

    def __pow__(self, e):
        if isinstance(e, EW):
            return Mat(self.a ** e.v)
        e = _num(e)
        if self.a.size == 1:
            return float(self.a.flat[0]) ** e
        if e == -1 and self.a.shape[0] == self.a.shape[1]:
            return Mat(np.linalg.inv(self.a))                   # MATLAB: inv via LU
        raise MlabError("matrix power other than ^-1 is not in the subset")

    def __float__(self): return _num(self)
    def __bool__(self): return bool(np.all(self.a != 0)) and self.a.size > 0
    def __eq__(self, o): return _num(self) == _num(o)
    def __ne__(self, o): return _num(self) != _num(o)
    def __lt__(self, o): return _num(self) < _num(o)
    def __le__(self, o): return _num(self) <= _num(o)
    def __gt__(self, o): return _num(self) > _num(o)
    def __ge__(self, o): return _num(self) >= _num(o)
    __hash__ = None


class Char(str):
    """char row vector: indexable with ``s(1)``, ``s(2:end-1)``."""

    def __call__(self, *idx):
        if len(idx) != 1:
            raise MlabError("char arrays take one subscript here")
        k = _index(idx[0], len(self))
        if isinstance(k, (int, np.integer)):
            return Char(self[k])
        return Char("".join(self[j] for j in k))

    @property
    def T(self):
        return self


class StrMat:
    """MATLAB string matrix (``readmatrix(...,'OutputType','string')``); ``None`` is ``<missing>``."""

    def __init__(self, rows):
        r = len(rows)
        c = max((len(x) for x in rows), default=0)
        self.a = np.empty((r, c), dtype=object)
        for i, row in enumerate(rows):
            for j in range(c):
                v = row[j] if j < len(row) else None
                self.a[i, j] = None if v is None else Char(v)

    def _get(self, idx):
        if len(idx) == 1:
            k = _index(idx[0], self.a.size)
            return self.a.ravel(order="F")[k]
        return self.a[_index(idx[0], self.a.shape[0]), _index(idx[1], self.a.shape[1])]

    def __call__(self, *idx):
        v = self._get(idx)
        if isinstance(v, np.ndarray):
            raise MlabError("string-matrix slices are not in the subset")
        return v

    brace = __call__


class EW:
    """Exponent of ``.^`` (element-wise power)."""

    def __init__(self, v):
        self.v = _num(v)

    def __rpow__(self, base):
        return float(base) ** self.v


class Cell:
    """Cell array (2-D object array)."""

    @property
    def T(self):
        out = Cell(0, 0)
        out.a = self.a.T.copy()
        return out

    def __init__(self, r: int, c: int):
        self.a = np.empty((r, c), dtype=object)
        for i in range(r):
            for j in range(c):
                self.a[i, j] = Mat(np.zeros((0, 0)))

    @staticmethod
    def of(rows: Sequence[Sequence]):
        out = Cell(len(rows), len(rows[0]) if rows else 0)
        for i, r in enumerate(rows):
            for j, v in enumerate(r):
                out.a[i, j] = v
        return out

    def _rc(self, idx):
        if len(idx) == 1:
            n = self.a.size
            k = _index(idx[0], n)
            r, c = np.unravel_index(k, self.a.shape, order="F")
            return r, c, False
        return _index(idx[0], self.a.shape[0]), _index(idx[1], self.a.shape[1]), True

    def brace(self, *idx):
        r, c, two = self._rc(idx)
        if isinstance(r, (int, np.integer)) and isinstance(c, (int, np.integer)):
            return self.a[r, c]
        if two:                                                # comma-separated list: used inside [ ]
            sub = self.a[np.ix_(np.atleast_1d(r), np.atleast_1d(c))]
            return CsList(list(sub.ravel(order="F")))
        return CsList([self.a[i, j] for i, j in zip(np.atleast_1d(r), np.atleast_1d(c))])

    def __call__(self, *idx):
        r, c, two = self._rc(idx)
        if two:
            sub = self.a[np.ix_(np.atleast_1d(r), np.atleast_1d(c))]
        else:
            sub = np.array([self.a[i, j] for i, j in zip(np.atleast_1d(r), np.atleast_1d(c))], dtype=object).reshape(-1, 1)
        out = Cell(0, 0)
        out.a = sub.copy()
        return out

    def set(self, idx, val):
        r, c, two = self._rc(idx)
        if isinstance(val, Cell):
            vals = list(val.a.ravel(order="F"))
        else:
            raise MlabError("()-assignment into a cell needs a cell on the right")
        if two:
            rr, cc = np.atleast_1d(r), np.atleast_1d(c)
            if len(vals) != rr.size * cc.size:
                raise MlabError("cell assignment size mismatch")
            nr, nc = max(self.a.shape[0], int(rr.max()) + 1), max(self.a.shape[1], int(cc.max()) + 1)
            if (nr, nc) != self.a.shape:                       # MATLAB grows the cell, new slots hold []
                big = Cell(nr, nc)
                big.a[:self.a.shape[0], :self.a.shape[1]] = self.a
                self.a = big.a
            k = 0
            for j in cc:
                for i in rr:
                    self.a[i, j] = vals[k]
                    k += 1
        else:
            rr, cc = np.atleast_1d(r), np.atleast_1d(c)
            if len(vals) != rr.size:
                raise MlabError("cell assignment size mismatch")
            for i, j, v in zip(rr, cc, vals):
                self.a[i, j] = v

    def _grow_for(self, idx):
        """c{end+1} = v / c{i,j} = v beyond the current size: MATLAB grows the cell."""
        if len(idx) == 1:
            i = idx[0]
            k = (i.at(self.a.size) if isinstance(i, _End) else int(round(float(i)))) - 1
            if k >= self.a.size:
                if self.a.size and self.a.shape[0] != 1:
                    raise MlabError("linear growth of a non-row cell")
                big = Cell(1, k + 1)
                big.a[0, :self.a.size] = self.a.ravel(order="F")
                self.a = big.a
            return
        need = []
        for i, n in zip(idx, self.a.shape):
            need.append(max(n, i.at(n) if isinstance(i, _End) else int(round(float(i)))))
        if tuple(need) != self.a.shape:
            big = Cell(*need)
            big.a[:self.a.shape[0], :self.a.shape[1]] = self.a
            self.a = big.a

    def brace_set(self, idx, val):
        # 'end' refers to the size BEFORE the assignment grows the cell
        dims = (self.a.size,) if len(idx) == 1 else self.a.shape
        idx = tuple(float(i.at(n)) if isinstance(i, _End) else i for i, n in zip(idx, dims))
        self._grow_for(idx)
        r, c, _ = self._rc(idx)
        self.a[r, c] = val

    def brace_ref(self, *idx):
        """``c{i,j}`` as the base of a nested assignment (``c{i,j}{end+1} = v``): an empty slot
        becomes a cell, as MATLAB does."""
        r, c, _ = self._rc(idx)
        v = self.a[r, c]
        if isinstance(v, Mat) and v.a.size == 0:
            v = Cell(0, 0)
            self.a[r, c] = v
        return v


class CsList:
    """Comma-separated list produced by ``c{i,a:b}`` (only meaningful inside ``[ ]``)."""

    def __init__(self, items):
        self.items = items


class Struct:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def copy(self):
        return Struct(**self.__dict__)


class StructArray:
    """1 x n struct array (``data.points``)."""

    def __init__(self, items: List[Struct]):
        self.items = items

    def __call__(self, i):
        if i is COLON:
            return self
        k = _index(i, len(self.items))
        if k == len(self.items):                               # data.points(i).x = ... creates element i
            self.items.append(Struct())
        return self.items[k]

    def __getattr__(self, name):                              # data.points.x -> comma-separated list
        if name == "items":
            raise AttributeError(name)
        return CsList([getattr(it, name) for it in self.items])


class StrCol(list):
    """n x 1 string array (``TIE``): ``TIE(i)`` is the i-th string."""

    def __call__(self, i):
        return Char(self[_index(i, len(self))])

    brace = __call__


# ------------------------------------------------------------------------------------ built-ins

def _size(x, d=None):
    if isinstance(x, (Mat, Cell, StrMat)):
        shp = x.a.shape
    elif isinstance(x, StructArray):
        shp = (1, len(x.items))
    elif isinstance(x, str):
        shp = (1, len(x))
    elif isinstance(x, (list, tuple)):                       # string column (TIE)
        shp = (len(x), 1)
    else:
        shp = (1, 1)
    if d is None:
        return Mat([float(shp[0]), float(shp[1])])
    return float(shp[int(_num(d)) - 1])


def _length(x):
    r, c = _size(x, 1), _size(x, 2)
    return 0.0 if r == 0 or c == 0 else max(r, c)


def _zeros(r, c=None):
    c = r if c is None else c
    return Mat(np.zeros((int(_num(r)), int(_num(c)))))


def _cell(r, c=None):
    c = r if c is None else c
    return Cell(int(_num(r)), int(_num(c)))


def _mat(rows):
    """Matrix / cell literal from rows of already evaluated elements."""
    rows = [[e for x in r for e in (x.items if isinstance(x, CsList) else [x])] for r in rows]
    rows = [r for r in rows if r]
    if not rows:
        return Mat(np.zeros((0, 0)))
    flat = [e for r in rows for e in r]
    if any(isinstance(e, Cell) for e in flat):
        blocks = []
        for r in rows:
            # [c, 'a', 'b'] : non-cell elements are wrapped; empty cells vanish
            parts = [(e if isinstance(e, Cell) else Cell.of([[e]])) for e in r]
            parts = [e.a for e in parts if e.a.size] or [np.empty((0, 0), dtype=object)]
            blocks.append(np.concatenate(parts, axis=1))
        blocks = [b for b in blocks if b.size] or [np.empty((0, 0), dtype=object)]
        out = Cell(0, 0)
        out.a = np.concatenate(blocks, axis=0)
        return out
    if all(isinstance(e, str) for e in flat):
        if len(rows) != 1:
            raise MlabError("multi-row char literal")
        return Char("".join(flat))
    if any(isinstance(e, str) for e in flat):
        raise MlabError("mixed char / numeric literal")

    def arr(e):
        return e.a if isinstance(e, Mat) else np.array([[float(e)]])

    blocks = []
    for r in rows:
        parts = [p for p in (arr(e) for e in r) if p.size] or [np.zeros((0, 0))]
        if len({p.shape[0] for p in parts}) != 1:
            raise MlabError("horizontal concatenation: row counts differ")
        blocks.append(np.concatenate(parts, axis=1))
    blocks = [b for b in blocks if b.size] or [np.zeros((0, 0))]
    if len({b.shape[1] for b in blocks}) != 1:
        raise MlabError("vertical concatenation: column counts differ")
    return Mat(np.concatenate(blocks, axis=0))


def _elementwise(fn):
    def f(x):
        if isinstance(x, Mat):
            return Mat(np.vectorize(fn, otypes=[float])(x.a)) if x.a.size else Mat(x.a.copy())
        return fn(float(x))
    return f


def _diag(x):
    if not isinstance(x, Mat):
        return Mat([[float(x)]])
    if 1 in x.a.shape:
        return Mat(np.diag(x.a.ravel()))
    return Mat(np.diag(x.a).reshape(-1, 1))


def _repmat(x, r, c):
    a = x.a if isinstance(x, Mat) else np.array([[float(x)]])
    return Mat(np.tile(a, (int(_num(r)), int(_num(c)))))


def _sum(x):
    if not isinstance(x, Mat):
        return float(x)
    if 1 in x.a.shape:
        # MATLAB sums vectors in order; use the same left-to-right order, not pairwise
        t = 0.0
        for v in x.a.ravel():
            t += float(v)
        return t
    return Mat(x.a.sum(axis=0, keepdims=True))


def _power(x, e):
    if isinstance(x, Mat) and x.a.size != 1:
        return Mat(x.a ** _num(e))
    return _num(x) ** _num(e)


def _unwrap1(x):
    """1 x 1 cell -> its content (strcmp(CNT(j,1), id) compares a 1x1 cell with a char)."""
    if isinstance(x, Cell) and x.a.size == 1:
        return x.a.flat[0]
    return x


def _strcmp(a, b):
    a, b = _unwrap1(a), _unwrap1(b)
    return isinstance(a, str) and isinstance(b, str) and str(a) == str(b)


def _strcat(*parts):
    return Char("".join(str(p) for p in parts))


def _num2str(v, prec=None):
    """num2str for scalars: '%d' for integers, else '%.Ng' with N = max(floor(log10|x|) + 5, 5) (<= 16);
    num2str(x, n) = '%.ng'."""
    if isinstance(v, str):
        return Char(v)
    v = _num(v)
    if prec is not None:
        return Char("%.*g" % (int(_num(prec)), v))
    if math.isfinite(v) and v == int(v):
        return Char("%d" % int(v))
    if not math.isfinite(v):
        return Char("NaN" if math.isnan(v) else ("Inf" if v > 0 else "-Inf"))
    n = min(max(int(math.floor(math.log10(abs(v)))) + 5, 5), 16)
    return Char("%.*g" % (n, v))


_CONV = re.compile(r"%(?:(\d+)\$)?([-+ 0#]*)(\d+)?(?:\.(\d+))?([sdifeEgGc%])")


def _fprintf_text(fmt, args):
    """MATLAB fprintf semantics on a char format: escapes, positional %N$, cyclic reuse of the format."""
    fmt = str(fmt).replace("\\n", "\n").replace("\\t", "\t").replace("\\\\", "\\")
    flat = []
    for a in args:
        if isinstance(a, CsList):
            flat += list(a.items)
        elif isinstance(a, Mat):
            flat += [float(x) for x in a.a.ravel(order="F")]
        else:
            flat.append(a)
    convs = [m for m in _CONV.finditer(fmt) if m.group(5) != "%"]
    if not convs:
        return fmt.replace("%%", "%")
    positional = any(m.group(1) for m in convs)
    per = max(int(m.group(1)) for m in convs) if positional else len(convs)

    def once(vals):
        out, pos, k = [], 0, 0
        for m in _CONV.finditer(fmt):
            out.append(fmt[pos:m.start()])
            pos = m.end()
            if m.group(5) == "%":
                out.append("%")
                continue
            idx = int(m.group(1)) - 1 if m.group(1) else k
            k += 1
            if idx >= len(vals):
                break                                            # MATLAB stops at the first conversion without data
            v = vals[idx]
            conv = "d" if m.group(5) == "i" else m.group(5)
            spec = "%" + m.group(2) + (m.group(3) or "") + ("." + m.group(4) if m.group(4) is not None else "") + conv
            if conv in "di" and not isinstance(v, str):
                spec = spec if float(v) == int(float(v)) else spec[:-1] + "e"
                v = int(float(v)) if float(v) == int(float(v)) else float(v)
            elif conv in "feEgG":
                v = float(v)
            elif conv in "sc":
                v = str(v) if isinstance(v, str) else _num2str(v)
            out.append(spec % v)
        else:
            out.append(fmt[pos:])
        return "".join(out)

    if not flat:
        return once([])
    return "".join(once(flat[i:i + per]) for i in range(0, len(flat), per))


class FileSink:
    """fopen/fprintf/fclose target that keeps the text."""

    def __init__(self):
        self.parts = []

    def text(self):
        return "".join(self.parts)


def _fprintf(fid, fmt, *args):
    if not isinstance(fid, FileSink):
        raise MlabError("fprintf needs a FileSink as its first argument here")
    fid.parts.append(_fprintf_text(fmt, args))
    return 0.0


def _str2double(x):
    """MATLAB str2double for the plain decimal forms the data files use; NaN otherwise."""
    x = _unwrap1(x)
    if not isinstance(x, str) or "_" in x:
        return math.nan
    try:
        return float(x)
    except ValueError:
        return math.nan


def _isnan(x):
    return isinstance(x, float) and math.isnan(x)


def _unique(x):
    if isinstance(x, Cell):
        vals = sorted({str(v) for v in x.a.ravel(order="F")})
        return Cell.of([[Char(v)] for v in vals])
    if isinstance(x, Mat):
        return Mat(np.unique(x.a).reshape(1, -1) if x.a.shape[0] == 1 else np.unique(x.a).reshape(-1, 1))
    raise MlabError("unique: unsupported argument")


def _mean(x):
    if not isinstance(x, Mat):
        return float(x)
    if 1 in x.a.shape:
        return _sum(x) / x.a.size
    return Mat(x.a.mean(axis=0, keepdims=True))


def _max(x):
    return float(np.max(x.a)) if isinstance(x, Mat) else float(x)


def _fieldnames(s):
    return Cell.of([[Char(k)] for k in s.__dict__])


def _struct2cell(s):
    return Cell.of([[Char(v) if isinstance(v, str) else v] for v in s.__dict__.values()])


def _sortrows(c, col):
    j = int(_num(col)) - 1
    order = sorted(range(c.a.shape[0]), key=lambda i: str(c.a[i, j]))     # stable, like sortrows
    out = Cell(0, 0)
    out.a = c.a[order, :].copy()
    return out


def _fileparts(path):
    d, base = os.path.split(str(path).rstrip("/"))
    stem, ext = os.path.splitext(base)
    return Char(d), Char(stem), Char(ext)


def _rmfield(s: Struct, name: str):
    out = s.copy()
    out.__dict__.pop(name, None)
    return out


def _isempty(x):
    return _length(x) == 0


BUILTINS = {
    "size": _size, "length": _length, "zeros": _zeros, "cell": _cell, "sqrt": _elementwise(math.sqrt),
    "sin": _elementwise(math.sin), "cos": _elementwise(math.cos), "tan": _elementwise(math.tan),
    "atan": _elementwise(math.atan), "atan2": lambda y, x: math.atan2(_num(y), _num(x)),
    "sec": _elementwise(lambda t: 1.0 / math.cos(t)), "abs": _elementwise(abs), "strcmp": _strcmp,
    "strcat": _strcat, "num2str": _num2str, "diag": _diag, "repmat": _repmat, "sum": _sum, "rmfield": _rmfield,
    "isempty": _isempty, "pi": lambda: math.pi, "str2double": _str2double, "isnan": _isnan, "unique": _unique,
    "char": lambda x: Char(_unwrap1(x)), "ismissing": lambda x: x is None, "fileparts": _fileparts,
    "mean": _mean, "max": _max, "fieldnames": _fieldnames, "struct2cell": _struct2cell, "sortrows": _sortrows,
    "fopen": lambda *a: FileSink(), "fclose": lambda *a: 0.0,
    "true": True, "false": False, "__chr": Char, "ischar": lambda x: isinstance(x, str),
    "isa": lambda x, cls: (str(cls) == "double" and isinstance(x, (float, int, Mat)) and not isinstance(x, bool)),
    "__mat": _mat, "__rng": Rng, "__COLON": COLON, "__END": END, "__power": _power, "__Cell": Cell,
    "__cellwrap": lambda v: Cell.of([list(v.items)]) if isinstance(v, CsList) else Cell.of([[v]]),
    "__ew": EW,
}
DROPPED = ("disp", "errordlg", "waitfor", "tic", "warning", "clear", "close", "figure")
for _name in DROPPED:                                           # `h = errordlg(...)`: the call survives as a no-op
    BUILTINS[_name] = lambda *a: 0.0
BUILTINS["fprintf"] = _fprintf

# ------------------------------------------------------------------------------------ tokeniser

_OPERAND_END = ("num", "id", "str", "close", "transpose")


class Tok:
    __slots__ = ("kind", "text", "sp")

    def __init__(self, kind, text, sp):
        self.kind, self.text, self.sp = kind, text, sp

    def __repr__(self):
        return f"{self.kind}:{self.text!r}"


_NUM = re.compile(r"(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)")
_ID = re.compile(r"[A-Za-z_]\w*")
_OPS = ("...", "==", "~=", "<=", ">=", "&&", "||", ".*", "./", ".^", ".'")


def tokenize(src: str) -> List[Tok]:
    """MATLAB source text -> tokens; comments removed; newlines kept as tokens."""
    toks: List[Tok] = []
    i, n = 0, len(src)
    sp = False
    while i < n:
        ch = src[i]
        if ch in " \t\r":
            sp = True
            i += 1
            continue
        if ch == "\n":
            toks.append(Tok("nl", "\n", sp))
            sp = False
            i += 1
            continue
        if ch == "%":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if ch == "'":
            prev = toks[-1] if toks else None
            if prev is not None and not sp and prev.kind in _OPERAND_END:
                toks.append(Tok("transpose", "'", False))
                i += 1
                continue
            j = i + 1
            buf = []
            while True:
                if j >= n or src[j] == "\n":
                    raise MlabError("unterminated string")
                if src[j] == "'":
                    if j + 1 < n and src[j + 1] == "'":
                        buf.append("'")
                        j += 2
                        continue
                    break
                buf.append(src[j])
                j += 1
            toks.append(Tok("str", "".join(buf), sp))
            sp = False
            i = j + 1
            continue
        if ch == '"':
            j = src.index('"', i + 1)
            toks.append(Tok("str", src[i + 1:j], sp))
            sp = False
            i = j + 1
            continue
        m = _NUM.match(src, i)
        if m and (ch.isdigit() or ch == "."and i + 1 < n and src[i + 1].isdigit()):
            toks.append(Tok("num", m.group(0), sp))
            sp = False
            i = m.end()
            continue
        m = _ID.match(src, i)
        if m:
            toks.append(Tok("id", m.group(0), sp))
            sp = False
            i = m.end()
            continue
        for op in _OPS:
            if src.startswith(op, i):
                if op == "...":                                   # continuation: swallow to end of line
                    while i < n and src[i] != "\n":
                        i += 1
                    i += 1
                    sp = True
                    break
                toks.append(Tok("op", op, sp))
                sp = False
                i += len(op)
                break
        else:
            kind = {"(": "open", "[": "open", "{": "open", ")": "close", "]": "close", "}": "close",
                    ",": "comma", ";": "semi", ":": "colon"}.get(ch, "op")
            if kind == "op" and ch not in "+-*/^<>=~&|.@":
                raise MlabError(f"unexpected character {ch!r}")
            toks.append(Tok(kind, ch, sp))
            sp = False
            i += 1
    return toks


def _match(toks: List[Tok], i: int) -> int:
    """index of the bracket closing toks[i]"""
    depth = 0
    for j in range(i, len(toks)):
        if toks[j].kind == "open":
            depth += 1
        elif toks[j].kind == "close":
            depth -= 1
            if depth == 0:
                return j
    raise MlabError("unbalanced brackets")


def _split(toks: List[Tok], kinds: Sequence[str]) -> List[List[Tok]]:
    """split on top-level tokens of the given kinds"""
    out, cur, depth = [], [], 0
    for t in toks:
        if t.kind == "open":
            depth += 1
        elif t.kind == "close":
            depth -= 1
        if depth == 0 and t.kind in kinds:
            out.append(cur)
            cur = []
        else:
            cur.append(t)
    out.append(cur)
    return out


# ------------------------------------------------------------------------------------ expressions

_PYOP = {"~=": "!=", "&&": " and ", "||": " or ", ".*": "*", "./": "/", "~": " not ", "&": " and ", "|": " or ",
         ".'": ".T"}


def _elements(toks: List[Tok]) -> List[List[Tok]]:
    """One row of a [ ] literal -> its elements (MATLAB's blank rule)."""
    out, cur, depth = [], [], 0
    for k, t in enumerate(toks):
        if depth == 0 and t.kind == "comma":
            out.append(cur)
            cur = []
            continue
        if depth == 0 and cur and t.sp:
            prev = cur[-1]
            prev_end = prev.kind in _OPERAND_END
            starts = t.kind in ("num", "id", "str") or (t.kind == "open")
            if t.kind == "op" and t.text in "+-" and prev_end:
                nxt = toks[k + 1] if k + 1 < len(toks) else None
                starts = nxt is not None and not nxt.sp         # '1 -x' -> two elements, '1 - x' -> one
            if prev_end and starts:
                out.append(cur)
                cur = []
        if t.kind == "open":
            depth += 1
        elif t.kind == "close":
            depth -= 1
        cur.append(t)
    out.append(cur)
    return [e for e in out if e]


def _arg(toks: List[Tok], indexing: bool) -> str:
    """One call / index argument (handles ':' ranges)."""
    parts = _split(toks, ("colon",))
    if len(parts) == 1:
        return expr(toks, indexing)
    if all(not p for p in parts) and len(parts) == 2:
        return "__COLON"
    if len(parts) in (2, 3) and all(parts):
        return "__rng(" + ", ".join(expr(p, indexing) for p in parts) + ")"
    raise MlabError("unsupported ':' expression")


def expr(toks: List[Tok], indexing: bool = False) -> str:
    """MATLAB expression tokens -> Python expression source."""
    out: List[str] = []
    i = 0
    prev = None
    while i < len(toks):
        t = toks[i]
        if t.kind == "nl":
            i += 1
            continue
        if t.kind == "num":
            out.append(t.text if any(c in t.text for c in ".eE") else t.text + ".0")
        elif t.kind == "str":
            out.append("__chr(" + repr(t.text) + ")")
        elif t.kind == "id":
            if t.text == "end" and indexing:
                out.append("__END")
            else:
                out.append("_m_" + t.text if t.text in _PYTHON_RESERVED else t.text)
        elif t.kind == "transpose":
            out.append(".T")
        elif t.kind == "op":
            if t.text in ("^", ".^"):
                # a^b binds tighter than unary minus on the left and takes a signed operand on the right
                j = i + 1
                rhs: List[Tok] = []
                while j < len(toks) and toks[j].kind == "op" and toks[j].text in "+-":
                    rhs.append(toks[j])
                    j += 1
                if j >= len(toks):
                    raise MlabError("dangling ^")
                if toks[j].kind == "open":
                    k = _match(toks, j)
                    rhs += toks[j:k + 1]
                    j = k + 1
                else:
                    rhs.append(toks[j])
                    j += 1
                    while j < len(toks) and toks[j].kind == "open" and not toks[j].sp and toks[j].text == "(":
                        k = _match(toks, j)                       # f(x) / v(i) as exponent
                        rhs += toks[j:k + 1]
                        j = k + 1
                out.append(("**__ew(" if t.text == ".^" else "**(") + expr(rhs, indexing) + ")")
                prev = toks[j - 1]
                i = j
                continue
            if t.text == "." and i + 1 < len(toks) and toks[i + 1].kind == "id":
                out.append(".")
            elif t.text == "=":
                raise MlabError("assignment inside an expression")
            else:
                out.append(_PYOP.get(t.text, t.text))
        elif t.kind == "open":
            j = _match(toks, i)
            inner = toks[i + 1:j]
            attached = prev is not None and not t.sp and prev.kind in ("id", "close")
            if t.text == "(":
                if prev is not None and prev.kind in ("id", "close"):
                    args = [a for a in _split(inner, ("comma",))]
                    args = [] if args == [[]] else args
                    out.append("(" + ", ".join(_arg(a, True) for a in args) + ")")
                else:
                    out.append("(" + expr(inner, indexing) + ")")
            elif t.text == "[":
                rows = _split(toks[i + 1:j], ("semi", "nl"))
                def element(e):
                    parts = _split(e, ("colon",))
                    if len(parts) > 1:                          # [a b:c] : a range as an element
                        return "__rng(" + ", ".join(expr(p_, indexing) for p_ in parts) + ").mat()"
                    return expr(e, indexing)
                out.append("__mat([" + ", ".join(
                    "[" + ", ".join(element(e) for e in _elements(r)) + "]" for r in rows if r) + "])")
            else:  # {
                if attached:
                    args = _split(inner, ("comma",))
                    out.append(".brace(" + ", ".join(_arg(a, True) for a in args) + ")")
                else:
                    out.append("__cellwrap(" + expr(inner, indexing) + ")")
            prev = toks[j]
            i = j + 1
            continue
        elif t.kind == "colon":
            raise MlabError("':' outside an index or for-range")
        elif t.kind == "comma":
            raise MlabError("unexpected ','")
        elif t.kind == "semi":
            raise MlabError("unexpected ';'")
        else:
            raise MlabError(f"unexpected token {t!r}")
        prev = t
        i += 1
    return "".join(_space(o) for o in out).strip()


def _space(s: str) -> str:
    return s if s.startswith((".", "(", "**")) or s in (")",) else " " + s


_PYTHON_RESERVED = {"lambda", "from", "import", "class", "def", "in", "is", "not", "and", "or", "pass", "global",
                    "with", "as", "assert", "del", "except", "finally", "raise", "try", "yield", "None", "True",
                    "False", "nonlocal", "async", "await", "print", "exec"}

# ------------------------------------------------------------------------------------ statements


def _logical_lines(toks: List[Tok]) -> List[List[Tok]]:
    """Split a token stream into statements: newline or ';' / ',' at bracket depth 0."""
    out, cur, depth = [], [], 0
    for t in toks:
        if t.kind == "open":
            depth += 1
        elif t.kind == "close":
            depth -= 1
        if depth == 0 and t.kind in ("nl", "semi"):
            if cur:
                out.append(cur)
            cur = []
            continue
        if depth > 0 and t.kind == "nl":
            cur.append(t)                                       # row separator inside [ ]
            continue
        cur.append(t)
    if cur:
        out.append(cur)
    return out


def _find_assign(toks: List[Tok]) -> int:
    depth = 0
    for k, t in enumerate(toks):
        if t.kind == "open":
            depth += 1
        elif t.kind == "close":
            depth -= 1
        elif depth == 0 and t.kind == "op" and t.text == "=":
            return k
    return -1


class Program:
    """Transpiled MATLAB: functions by name plus helpers to run statement ranges of a script."""

    def __init__(self):
        self.env: Dict[str, object] = dict(BUILTINS)
        self.sources: Dict[str, str] = {}

    # ---- statements -> python lines
    def _emit(self, stmts: List[List[Tok]], outs: Optional[List[str]], base_indent: int,
              ret: Optional[str] = None) -> List[str]:
        lines: List[str] = []
        ind = base_indent
        stack: List[str] = []

        def put(s):
            lines.append("    " * ind + s)

        if ret is None:
            ret = "return " + (("(" + ", ".join(outs) + ",)") if outs and len(outs) > 1 else (outs[0] if outs else "None"))
        for st in stmts:
            st = [t for t in st if not (t.kind == "comma" and False)]
            head = st[0]
            if head.kind == "id" and head.text in DROPPED:
                continue
            if head.kind == "id" and head.text in ("if", "while", "elseif"):
                cond = expr(st[1:])
                if head.text == "elseif":
                    ind -= 1
                    put(f"elif {cond}:")
                else:
                    put(f"{head.text} {cond}:")
                    stack.append(head.text)
                ind += 1
                put("pass")
                continue
            if head.kind == "id" and head.text == "else":
                ind -= 1
                put("else:")
                ind += 1
                put("pass")
                continue
            if head.kind == "id" and head.text == "for":
                k = _find_assign(st)
                var = st[1].text
                rng = _split(st[k + 1:], ("colon",))
                put(f"for {var} in __rng({', '.join(expr(p) for p in rng)}):")
                stack.append("for")
                ind += 1
                put("pass")
                continue
            if head.kind == "id" and head.text == "end" and len(st) == 1:
                if not stack:
                    raise MlabError("'end' without a block")
                stack.pop()
                ind -= 1
                continue
            if head.kind == "id" and head.text in ("break", "continue") and len(st) == 1:
                put(head.text)
                continue
            if head.kind == "id" and head.text == "return" and len(st) == 1:
                put(ret)
                continue
            k = _find_assign(st)
            if k < 0:
                # expression statement: a bare call keeps its side effects, a bare value is MATLAB's echo
                put(expr(st))
                continue
            lhs, rhs = st[:k], st[k + 1:]
            if lhs[0].kind == "open" and lhs[0].text == "[":        # [a, b] = f(...)
                names = [e for e in _elements(lhs[1:_match(lhs, 0)])]
                tgt = ", ".join("_" if (len(e) == 1 and e[0].text == "~") else expr(e) for e in names)
                put(f"({tgt},) = {expr(rhs)}")
                continue
            # indexed assignment?  name(...) = / name{...} = / a.b(...) = ...
            last_open = None
            depth = 0
            for j, t in enumerate(lhs):
                if t.kind == "open":
                    if depth == 0:
                        last_open = j
                    depth += 1
                elif t.kind == "close":
                    depth -= 1
            if last_open is not None and _match(lhs, last_open) == len(lhs) - 1:
                base = expr(lhs[:last_open])
                args = _split(lhs[last_open + 1:-1], ("comma",))
                a = ", ".join(_arg(x, True) for x in args)
                meth = "brace_set" if lhs[last_open].text == "{" else "set"
                if meth == "brace_set" and ".brace(" in base and base.rstrip().endswith(")"):
                    k2 = base.rindex(".brace(")                # c{i,j}{end+1} = v : the inner slot becomes a cell
                    base = base[:k2] + ".brace_ref(" + base[k2 + len(".brace("):]
                put(f"{base}.{meth}(({a},), {expr(rhs)})")
                continue
            put(f"{expr(lhs)} = {expr(rhs)}")
        if stack:
            raise MlabError("unterminated block")
        return lines

    def add_functions(self, src: str, only: Optional[Sequence[str]] = None) -> List[str]:
        """Define every ``function`` of a source file (optionally only the named ones)."""
        stmts = _logical_lines(tokenize(src))
        names = []
        k = 0
        while k < len(stmts):
            st = stmts[k]
            if not (st[0].kind == "id" and st[0].text == "function"):
                k += 1
                continue
            eq = _find_assign(st)
            if eq >= 0:
                lhs = st[1:eq]
                outs = ([e[0].text for e in _elements(lhs[1:-1])] if lhs[0].kind == "open" else [lhs[0].text])
                sig = st[eq + 1:]
            else:
                outs, sig = [], st[1:]
            name = sig[0].text
            args = [a[0].text for a in _split(sig[2:-1], ("comma",)) if a] if len(sig) > 1 else []
            # body: up to the matching 'end' (block depth) or the next 'function'
            depth, j = 1, k + 1
            body = []
            while j < len(stmts):
                h = stmts[j][0]
                if h.kind == "id" and h.text in ("if", "for", "while", "switch", "try"):
                    depth += 1
                elif h.kind == "id" and h.text == "end" and len(stmts[j]) == 1:
                    depth -= 1
                    if depth == 0:
                        break
                elif h.kind == "id" and h.text == "function":
                    break
                body.append(stmts[j])
                j += 1
            k = j + 1 if (j < len(stmts) and stmts[j][0].text == "end") else j
            if only is not None and name not in only:
                continue
            ret = "return " + (("(" + ", ".join(outs) + ",)") if len(outs) > 1 else (outs[0] if outs else "None"))
            lines = [f"def {name}({', '.join(a + '=None' for a in args)}):",
                     f"    nargin = float(len([__a for __a in ({' '.join(a + ',' for a in args)}) if __a is not None]))"] \
                + self._emit(body, outs, 1) + ["    " + ret]
            code = "\n".join(lines)
            self.sources[name] = code
            exec(compile(code, f"<mlab:{name}>", "exec"), self.env)
            names.append(name)
        return names

    def run(self, src: str, variables: Dict[str, object], label: str = "script") -> Dict[str, object]:
        """Execute script statements with the given workspace; returns the workspace afterwards.
        ``return`` ends the script."""
        stmts = _logical_lines(tokenize(src))
        body = self._emit(stmts, None, 1, ret="return dict(locals())")
        names = sorted(variables)
        code = "\n".join([f"def __script({', '.join(names)}):"] + body + ["    return dict(locals())"])
        self.sources[label] = code
        exec(compile(code, f"<mlab:{label}>", "exec"), self.env)
        ws = self.env["__script"](**variables)
        ws = {k: v for k, v in ws.items() if not k.startswith("__")}
        return ws


def read_lines(relpath: str, first: int, last: int) -> str:
    """Lines first..last (1-based, inclusive) of a file of the reference tree."""
    with open(os.path.join(REFERENCE_ROOT, relpath), "r") as fh:
        lines = fh.read().split("\n")
    return "\n".join(lines[first - 1:last]) + "\n"


def read_file(relpath: str) -> str:
    with open(os.path.join(REFERENCE_ROOT, relpath), "r") as fh:
        return fh.read()
