"""Text formats of the reference: .pho/.ext/.cnt/.int/.tie/.cze/.cfg.

Host-side mirror of the reference's I/O layer (it stays host code; nothing here is
on the device path):

* ``ReadFiles``   -- reference ``functions/ReadFiles.m:4-52``: one file per extension is
  located by globbing a directory (``ReadFiles.m:24``) and read as a string table with
  ``readmatrix(..., 'Delimiter',{' ','\\t'}, 'ConsecutiveDelimitersRule','join',
  'LeadingDelimitersRule','ignore', 'OutputType','string', 'CommentStyle','#')``
  (``ReadFiles.m:49``).  The GUI dialogs of ``ReadFiles.m:25-44`` (more than one / no
  file) become errors (``terminate = 1``).
* ``findSetting`` -- reference ``functions/findSetting.m:7-55``.
* writers for the same formats so synthetic networks can be consumed by ``main.m``.
"""
from __future__ import annotations

import glob
import math
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

StringTable = List[List[Optional[str]]]


def read_string_table(path: str) -> StringTable:
    """``readmatrix(...,'OutputType','string','CommentStyle','#')`` (ReadFiles.m:49).

    Rows are the non-empty lines after comment stripping; delimiters are blanks and
    tabs, runs are joined and leading ones ignored.  Short rows are padded with
    ``None`` (MATLAB ``<missing>``) up to the widest row.
    """
    rows: StringTable = []
    with open(path, "r") as fh:
        for line in fh:
            hash_pos = line.find("#")
            if hash_pos >= 0:
                line = line[:hash_pos]
            toks = line.replace("\t", " ").split()
            if toks:
                rows.append(list(toks))
    width = max((len(r) for r in rows), default=0)
    for r in rows:
        r.extend([None] * (width - len(r)))
    return rows


def ReadFiles(exts: Sequence[str], folder: str = ".") -> Tuple[int, List[StringTable]]:
    """``[terminate, filecontents] = ReadFiles(files)`` (ReadFiles.m:4).

    Exactly one ``*<ext>`` must exist in ``folder`` for each requested extension;
    otherwise ``terminate`` is 1 (the reference opens a dialog there,
    ReadFiles.m:25-44) and the contents gathered so far are returned.
    """
    contents: List[StringTable] = []
    terminate = 0
    for ext in exts:
        hits = sorted(glob.glob(os.path.join(folder, "*" + ext)))
        if len(hits) != 1:
            print(f"Error on {ext}")
            terminate = 1
            break
        contents.append(read_string_table(hits[0]))
    return terminate, contents


def str2double(tok: Optional[str]) -> float:
    """MATLAB ``str2double``: NaN when the text is not a number."""
    if tok is None:
        return math.nan
    try:
        return float(tok)
    except ValueError:
        return math.nan


def findSetting(CFG: StringTable, name: str, error_tally: int, check01: bool = False):
    """``[value,error_tally] = findSetting(CFG,str,error_tally,Check_truefalse)``.

    findSetting.m:15-30 first match wins; a value wrapped in single quotes is text,
    anything else goes through ``str2double``.  Missing (``:33-37``), NaN (``:40-44``)
    and not-0/1 (``:47-53``) each add one to the tally.
    """
    value = -1
    found = False
    for row in CFG:
        if row and row[0] == name:
            found = True
            sval = row[1] if len(row) > 1 and row[1] is not None else ""
            if len(sval) >= 1 and sval[0] == "'" and sval[-1] == "'":
                value = sval[1:-1]
            else:
                value = str2double(sval)
            break
    if not found:
        print(f"warning:findSetting() could not find setting {name}")
        return value, error_tally + 1
    if isinstance(value, float) and math.isnan(value):
        print(f"Error:findSetting() {name} cannot be NaN.")
        return value, error_tally + 1
    if check01 and value != 1 and value != 0:
        print(f"Error:findSetting() {name} must be 1 or 0")
        return value, error_tally + 1
    return value, error_tally


# --------------------------------------------------------------------------- writers


def _fmt(v: float) -> str:
    return repr(float(v))


def write_pho(path, point_ids, image_ids, x, y):
    """.pho: pointID imageID x y (main.m:52, main.m:199-204)."""
    with open(path, "w") as fh:
        for p, i, a, b in zip(point_ids, image_ids, x, y):
            fh.write(f"{p}\t{i}\t{_fmt(a)}\t{_fmt(b)}\n")


def write_ext(path, image_ids, camera_ids, eop_rad):
    """.ext: imageID cameraID Xc Yc Zc omega phi kappa; angles in DEGREES (main.m:215-217)."""
    with open(path, "w") as fh:
        for i, c, e in zip(image_ids, camera_ids, np.asarray(eop_rad, dtype=np.float64)):
            deg = [math.degrees(a) for a in e[3:6]]
            fh.write(f"{i}\t{c}\t{_fmt(e[0])}\t{_fmt(e[1])}\t{_fmt(e[2])}\t"
                     f"{_fmt(deg[0])}\t{_fmt(deg[1])}\t{_fmt(deg[2])}\n")


def write_cnt(path, point_ids, xyz):
    """.cnt / .cze: pointID X Y Z (main.m:222-227)."""
    with open(path, "w") as fh:
        for p, r in zip(point_ids, np.asarray(xyz, dtype=np.float64)):
            fh.write(f"{p}\t{_fmt(r[0])}\t{_fmt(r[1])}\t{_fmt(r[2])}\n")


def write_int(path, camera_ids, cam_box, iop):
    """.int: two rows per camera (main.m:231-256).

    row 1: cameraID y_dir xmin ymin xmax ymax;  row 2: xp yp c k1..kNK p1 p2
    """
    with open(path, "w") as fh:
        for c, box, row in zip(camera_ids, np.asarray(cam_box), np.asarray(iop)):
            fh.write(f"{c}\t{int(box[0])}\t" + "\t".join(_fmt(v) for v in box[1:5]) + "\n")
            fh.write("\t".join(_fmt(v) for v in row) + "\n")


def write_tie(path, point_ids):
    with open(path, "w") as fh:
        for p in point_ids:
            fh.write(f"{p}\n")


def write_cfg(path, settings: dict):
    """config.cfg: ``name value`` rows; text values in single quotes (config.cfg:1-43)."""
    with open(path, "w") as fh:
        fh.write("# generated settings file (same grammar as the reference's config.cfg)\n")
        for k, v in settings.items():
            if isinstance(v, str):
                fh.write(f"{k}\t'{v}'\n")
            elif isinstance(v, float) and not float(v).is_integer():
                fh.write(f"{k}\t{_fmt(v)}\n")
            else:
                fh.write(f"{k}\t{int(v) if float(v).is_integer() else _fmt(v)}\n")
