"""GPU check of the block-sparse form of the reduced system (FEBA_SPARSE=1, csrc/feba_sparse.h), run as a
separate process by tests/test_zz_sparse_gpu.py (the environment switch is read when a handle is created, and a
fault in this opt-in path must not take the rest of the GPU suite with it).

For a free network and a network with control points: one step and the whole adjustment with the masked
supertile factorisation + sparse datum against (a) the dense form of the same library and (b) the oracle.
Tolerances: step 1e-7 relative (the dense forms of the library and of the oracle differ by that much from
extended precision on these networks, oracle/exact.py), end to end v 1e-8 max|v|, sigma02 1e-8, same iteration count.

    python tests/sparse_gpu_check.py [n_img] [n_pts] [tile_blocks]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(prob, x0, sparse_on, tile):
    import feba_b200 as fb
    os.environ["FEBA_SPARSE"] = "1" if sparse_on else "0"
    os.environ["FEBA_DAG_TILE"] = str(tile)
    with fb.Handle(prob) as h:
        info = h.sparse_info()
        h.set_xhat(x0)
        h.iterate()
        d1 = h.get_delta().copy()
        h.set_xhat(x0)
        it, trace = h.solve()
        res = h.residuals()
        xh = h.get_xhat().copy()
    return dict(info=info, d1=d1, it=it, v=res["v"], sigma02=res["sigma02"], xhat=xh)


def main():
    import feba_b200 as fb
    from oracle import sparse
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
    tile = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ok = True
    for mode, kw in (("free", {}), ("mixed", {"n_control": 60})):
        prob = fb.synth.make_network(n_img, n_pts, 8, 811, mode=mode, **kw)
        err, x0, _ = fb.Buildxhat(prob)
        dense = run(prob, x0, False, tile)
        sp = run(prob, x0, True, tile)
        ref = sparse.gauss_newton(prob, x0)
        assert not dense["info"]["active"]
        print(f"[{mode}] u_c {prob.u_c}  sparse {sp['info']}")
        if not sp["info"]["active"]:
            print(f"[{mode}] FAIL: FEBA_SPARSE=1 did not activate (reduced system too small for the task graph?)")
            ok = False
            continue
        e_step = np.linalg.norm(sp["d1"] - dense["d1"]) / np.linalg.norm(dense["d1"])
        e_v = np.max(np.abs(sp["v"] - ref["v"])) / np.max(np.abs(ref["v"]))
        e_s = abs(sp["sigma02"] - ref["sigma02"]) / ref["sigma02"]
        e_x = np.max(np.abs(sp["xhat"] - dense["xhat"]) / (np.abs(dense["xhat"]) + 1.0))
        print(f"[{mode}] step vs dense form {e_step:.2e}   iterations {sp['it']} (dense {dense['it']}, oracle "
              f"{ref['iterations']})   v vs oracle {e_v:.2e}   sigma02 {e_s:.2e}   xhat vs dense form {e_x:.2e}")
        good = (e_step < 1e-7 and sp["it"] == ref["iterations"] and e_v < 1e-8 and e_s < 1e-8)
        if mode == "free":
            good = good and sp["info"]["datum_images"] == 8
        if not good:
            print(f"[{mode}] FAIL")
            ok = False
    print("sparse form ok" if ok else "sparse form FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
