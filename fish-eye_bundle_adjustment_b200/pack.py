"""ctypes binding of ``include/feba_pack.h``: the native (C++, multi-threaded, hashed) problem build.

``load_problem_native(folder)`` returns the same :class:`Problem` as ``problem.load_problem`` --
the interpreted mirror of main.m:60-384 that the tests keep as the statement of what the files mean
-- but tokenises the .pho file and resolves the image / target / tie IDs in ``libfeba.so``
(at 10M observations the reference's O(n_obs*(nImg+nPts)) ``strcmp`` scans,
main.m:286-375, dominate the wall time, not the iteration).  The settings file is still read by
``findSetting`` (a few dozen rows).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import List, Optional

import numpy as np

from . import formats
from .problem import Problem, settings_from_cfg

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)

PACK_EXPORTS = ("feba_pack_read", "feba_pack_get", "feba_pack_problem", "feba_pack_ids", "feba_pack_timing",
                "feba_pack_last_error", "feba_pack_free")


class PackView(C.Structure):
    _fields_ = [("n_obs", C.c_int64), ("n_img", C.c_int32), ("n_cam", C.c_int32), ("n_pts", C.c_int32),
                ("n_tie", C.c_int32), ("n_iop_cols", C.c_int32), ("reserved", C.c_int32),
                ("obs_x", _pd), ("obs_y", _pd), ("obs_img", _pi), ("obs_pt", _pi), ("img_cam", _pi),
                ("eop0", _pd), ("iop0", _pd), ("cam_box", _pd), ("xyz0", _pd), ("pt_tie", _pi), ("tie_pt", _pi)]


class PackError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(text)
        self.code, self.text = code, text


_bound = False


def _lib():
    global _bound
    from .lib import load
    lib = load()
    if not _bound:
        P = C.c_void_p
        lib.feba_pack_read.argtypes = [C.c_char_p] * 5 + [C.c_int32, C.c_int32, C.c_int32, C.POINTER(P)]
        lib.feba_pack_get.argtypes = [P, C.POINTER(PackView)]
        lib.feba_pack_problem.argtypes = [P, C.c_void_p, C.c_void_p]
        lib.feba_pack_ids.argtypes = [P, C.c_int32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.feba_pack_timing.argtypes = [P, _pd]
        lib.feba_pack_last_error.argtypes = []
        lib.feba_pack_last_error.restype = C.c_char_p
        lib.feba_pack_free.argtypes = [P]
        lib.feba_pack_free.restype = None
        _bound = True
    return lib


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True).reshape(shape)


def _ids(lib, pack, which: int) -> List[str]:
    need = C.c_size_t(0)
    lib.feba_pack_ids(pack, which, None, 0, C.byref(need))
    if need.value == 0:
        return []
    buf = C.create_string_buffer(need.value)
    lib.feba_pack_ids(pack, which, buf, need.value, C.byref(need))
    return buf.raw[:need.value].decode().split("\n")[:-1]


def pack_files(pho: str, ext: str, cnt: str, intr: str, tie: Optional[str], num_radial: int, all_gcp: bool,
               threads: int = 0) -> dict:
    """feba_pack_read + copies of every array (the pack is released before returning)."""
    lib = _lib()
    pack = C.c_void_p()
    enc = lambda s: None if s is None else os.fsencode(s)
    rc = lib.feba_pack_read(enc(pho), enc(ext), enc(cnt), enc(intr), enc(tie), int(num_radial), int(bool(all_gcp)),
                            int(threads), C.byref(pack))
    if rc != 0:
        raise PackError(rc, (lib.feba_pack_last_error() or b"").decode())
    try:
        v = PackView()
        lib.feba_pack_get(pack, C.byref(v))
        sec = (C.c_double * 3)()
        lib.feba_pack_timing(pack, sec)
        out = dict(
            obs_x=_arr(v.obs_x, (v.n_obs,), np.float64), obs_y=_arr(v.obs_y, (v.n_obs,), np.float64),
            obs_img=_arr(v.obs_img, (v.n_obs,), np.int32), obs_pt=_arr(v.obs_pt, (v.n_obs,), np.int32),
            img_cam=_arr(v.img_cam, (v.n_img,), np.int32), eop0=_arr(v.eop0, (v.n_img, 6), np.float64),
            iop0=_arr(v.iop0, (v.n_cam, v.n_iop_cols), np.float64), cam_box=_arr(v.cam_box, (v.n_cam, 5), np.float64),
            xyz0=_arr(v.xyz0, (v.n_pts, 3), np.float64), pt_tie=_arr(v.pt_tie, (v.n_pts,), np.int32),
            tie_pt=_arr(v.tie_pt, (v.n_tie,), np.int32),
            point_ids=_ids(lib, pack, 0), image_ids=_ids(lib, pack, 1), camera_ids=_ids(lib, pack, 2),
            tie_ids=_ids(lib, pack, 3), seconds=tuple(sec))
    finally:
        lib.feba_pack_free(pack)
    return out


def _one(folder: str, ext: str) -> Optional[str]:
    hits = sorted(glob.glob(os.path.join(folder, "*" + ext)))       # ReadFiles.m:24
    if len(hits) != 1:
        print(f"Error on {ext}")                                     # ReadFiles.m:25-44 opens a dialog here
        return None
    return hits[0]


def load_problem_native(folder: str, cfg_folder: Optional[str] = None, threads: int = 0) -> Optional[Problem]:
    """main.m:60-384 for a data folder, the file parsing and ID resolution done by ``feba_pack_read``.
    Same return convention as ``problem.load_problem`` (None where main.m sets ``main_error``)."""
    cfg_dir = folder
    if not any(f.endswith(".cfg") for f in os.listdir(folder)):      # main.m:66-85
        cfg_dir = cfg_folder if cfg_folder is not None else folder
    term, files = formats.ReadFiles([".cfg"], cfg_dir)
    if term:
        print("Error reading files")
        return None
    s = settings_from_cfg(files[0], os.path.basename(os.path.abspath(folder)))
    if s is None:
        return None
    paths = [_one(folder, e) for e in (".pho", ".ext", ".cnt", ".int")]
    if any(p is None for p in paths):
        print("Error reading files")
        return None
    tie = None
    if s.Estimate_tie == 1 and s.Estimate_AllGCP == 0:               # main.m:180-188
        tie = _one(folder, ".tie")
        if tie is None:
            print("Error reading files")
            return None
    try:
        a = pack_files(*paths, tie, s.Num_Radial_Distortions, s.Estimate_AllGCP == 1, threads)
    except PackError as exc:
        print(exc.text)
        return None
    if s.Estimate_AllGCP == 1:                                       # main.m:260-264
        s.Estimate_tie = 1
    prob = Problem(settings=s, obs_x=a["obs_x"], obs_y=a["obs_y"], obs_img=a["obs_img"], obs_pt=a["obs_pt"],
                   img_cam=a["img_cam"], eop0=a["eop0"], iop0=a["iop0"], cam_box=a["cam_box"], xyz0=a["xyz0"],
                   pt_tie=a["pt_tie"], tie_pt=a["tie_pt"], point_ids=a["point_ids"], image_ids=a["image_ids"],
                   camera_ids=a["camera_ids"])
    prob.pack_seconds = a["seconds"]
    return prob
