"""Supertile pattern of the reduced camera system of a BASELINE configuration, counted with the library's own
pattern code (csrc/feba_sparse.h through the host-compiled test library): non-zero lower supertiles with
symbolic fill, flop and task count of the masked factorisation by supertile size.  CPU only.

    python scripts/sparse_pattern_stats.py [config index 1..4, default 3] [scale]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feba_b200 as fb                                   # noqa: E402

SRC = os.path.join(ROOT, "tests", "host_model", "model_host.cpp")
LIB = os.path.join(ROOT, "tests", "_build", "libfeba_model_host.so")
_pi = C.POINTER(C.c_int)


def host():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                    "-Wno-unknown-pragmas", SRC, "-o", LIB], check=True)
    lib = C.CDLL(LIB)
    lib.feba_host_sparse_datum.argtypes = [C.c_int] * 4 + [_pi]
    lib.feba_host_sparse_pattern.argtypes = [C.c_int] * 7 + [_pi, C.c_int, _pi, C.POINTER(C.c_ubyte)]
    return lib


def main():
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lib = host()
    prob = fb.synth.baseline_config(idx, scale=scale)
    n_img, ui = prob.numImg, prob.settings.u_perimage
    tie = prob.pt_tie[prob.obs_pt] >= 0
    A = sp.csr_matrix((np.ones(int(tie.sum()), dtype=np.float32), (prob.obs_pt[tie], prob.obs_img[tie])),
                      shape=(prob.numPts, n_img))
    Bl = sp.tril((A.T @ A).tocsr(), k=-1).tocoo()
    blocks = np.ascontiguousarray(np.stack([Bl.row, Bl.col], 1).astype(np.int32))
    off_cam, n_red = ui * n_img, prob.u_c
    nb = (n_red + 63) // 64
    print(f"configs[{idx}] scale {scale}: {n_img} images, {prob.n_obs} observations, u_c {n_red} ({nb} blocks); "
          f"{len(blocks) + n_img} of {n_img * (n_img + 1) // 2} lower image blocks non-zero "
          f"({100.0 * (len(blocks) + n_img) / (n_img * (n_img + 1) // 2):.2f} %)")
    dense = (64.0 * nb) ** 3 / 3
    print("| T | supertiles | non-zero lower (with fill) | flop | vs dense | tasks |\n|---|---|---|---|---|---|")
    for T in (4, 6, 8, 10, 12, 14, 16, 20, 24):
        if nb < 2 * T:
            continue
        datum = np.zeros(8, np.int32)
        nd = lib.feba_host_sparse_datum(n_img, ui, nb, T, datum.ctypes.data_as(_pi)) if prob.settings.Inner_Constraints else 0
        NT = (nb + T - 1) // T
        nz = np.zeros((NT + 1, NT + 1), np.uint8)
        lib.feba_host_sparse_pattern(nb, T, ui, n_img, off_cam, n_red, len(blocks), blocks.ctypes.data_as(_pi), nd,
                                     datum.ctypes.data_as(_pi), nz.ctypes.data_as(C.POINTER(C.c_ubyte)))
        nzb = nz.astype(bool)
        size = lambda t: ((nb - T * (NT - 1)) if t == NT - 1 else T) * 64 if t < NT else 64
        fl, tasks = 0.0, 0
        for k in range(NT):
            nk = size(k)
            fl += nk ** 3 / 3
            rows = [i for i in range(k + 1, NT + 1) if nzb[i, k]]
            for i in rows:
                fl += size(i) * nk * nk
                tasks += 1
                for j in rows:
                    if j > i or (j == NT and i != NT):
                        continue
                    fl += (2 if i != j else 1) * size(i) * size(j) * nk
                    tasks += 1
        low = np.tril(np.ones((NT, NT), bool))
        print(f"| {T} | {NT} | {nzb[:NT, :NT][low].sum()} / {low.sum()} | {fl:.2e} | {dense / fl:.1f}x fewer | {tasks + NT} |")


if __name__ == "__main__":
    main()
