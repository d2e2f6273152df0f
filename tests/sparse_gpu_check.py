"""GPU check of the nested-dissection plan of the reduced system (csrc/feba_order.h, feba_sparse.h), run as a
separate process by tests/test_zz_sparse_gpu.py (plan parameters are read from the environment when a handle is
created).

For a free network and a network with control points: one step and the whole adjustment with the plan's masked
supertile factorisation (+ sparse datum on the free network) against (a) the dense form of the same library
(plan = -1) and (b) an oracle -- oracle/sparse.py below 400 images, the C restatement oracle/cport.py above.
Tolerances: step 1e-7 relative (the dense forms of the library and of the oracle differ by that much from
extended precision on these networks, oracle/exact.py), end to end xhat 1e-9 group-normalised, v 1e-8 max|v|,
sigma02 1e-8, same iteration count.

    python tests/sparse_gpu_check.py [n_img] [n_pts] [leaf_images] [tile_max] [plan: 1 force | 0 automatic]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(prob, x0, plan):
    import feba_b200 as fb
    with fb.Handle(prob, plan=plan) as h:
        info, pinfo = h.sparse_info(), h.plan_info()
        h.set_xhat(x0)
        h.iterate()
        d1 = h.get_delta().copy()
        h.set_xhat(x0)
        it, trace = h.solve()
        res = h.residuals()
        xh = h.get_xhat().copy()
    return dict(info=info, plan=pinfo, d1=d1, it=it, v=res["v"], sigma02=res["sigma02"], xhat=xh)


def main():
    import feba_b200 as fb
    from oracle import cport, sparse
    from oracle.compare import group_rel
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
    os.environ["FEBA_ND_LEAF"] = sys.argv[3] if len(sys.argv) > 3 else "12"
    os.environ["FEBA_TILE_MAX"] = sys.argv[4] if len(sys.argv) > 4 else "2"
    plan = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    ok = True
    for mode, kw in (("free", {}), ("mixed", {"n_control": 60})):
        prob = fb.synth.make_network(n_img, n_pts, 8, 811, mode=mode, **kw)
        err, x0, _ = fb.Buildxhat(prob)
        dense = run(prob, x0, -1)
        sp = run(prob, x0, plan)
        sp2 = run(prob, x0, plan)
        # the task graph only reorders INDEPENDENT tiles: reruns are bit-identical (a missing dependency -- a race --
        # shows up here; compute-sanitizer is not available on this pool)
        if not (np.array_equal(sp["xhat"], sp2["xhat"]) and np.array_equal(sp["v"], sp2["v"]) and np.array_equal(sp["d1"], sp2["d1"])):
            print(f"[{mode}] FAIL: two runs of the plan form differ")
            ok = False
        ref = sparse.gauss_newton(prob, x0) if n_img < 400 else cport.CPort(prob).gauss_newton(x0)
        assert not dense["info"]["active"]
        print(f"[{mode}] u_c {prob.u_c}  plan {sp['plan']}  {sp['info']}")
        if not sp["info"]["active"]:
            print(f"[{mode}] FAIL: the nested-dissection plan did not activate")
            ok = False
            continue
        e_step = np.linalg.norm(sp["d1"] - dense["d1"]) / np.linalg.norm(dense["d1"])
        e_v = np.max(np.abs(sp["v"] - ref["v"])) / np.max(np.abs(ref["v"]))
        e_s = abs(sp["sigma02"] - ref["sigma02"]) / ref["sigma02"]
        e_x = group_rel(prob, sp["xhat"], dense["xhat"])
        e_xo = group_rel(prob, sp["xhat"], ref["xhat"])
        e_do = group_rel(prob, dense["xhat"], ref["xhat"])
        print(f"[{mode}] step vs dense form {e_step:.2e}   iterations {sp['it']} (dense {dense['it']}, oracle "
              f"{ref['iterations']})   v vs oracle {e_v:.2e}   sigma02 {e_s:.2e}   xhat vs dense form {e_x:.2e}   "
              f"xhat vs oracle {e_xo:.2e} (dense form vs oracle {e_do:.2e})")
        good = (e_step < 1e-7 and sp["it"] == ref["iterations"] == dense["it"] and e_v < 1e-8 and e_s < 1e-8
                and e_xo < 1e-9 and e_do < 1e-9)
        if mode == "free":
            good = good and sp["info"]["datum_images"] == 8
        if sp["plan"]["chain_blocks"] >= sp["plan"]["rows"] // 64:
            print(f"[{mode}] FAIL: the block was not cut")
            good = False
        if not good:
            print(f"[{mode}] FAIL")
            ok = False
    print("plan form ok" if ok else "plan form FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
