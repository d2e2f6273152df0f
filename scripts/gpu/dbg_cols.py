import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
sys.path.insert(0, os.getcwd())
import feba_b200 as fb
from oracle import sparse
prob = fb.synth.make_network(260, 20000, 8, 77, mode="mixed", n_control=50)
err, x0, _ = fb.Buildxhat(prob)
nb = sparse.normal_blocks(prob, x0)
d_ref, S_ref, g_ref = sparse.reduce_and_solve(prob, nb)
d_ref = sparse.unscale(prob, nb["L"], nb["q"], d_ref)
print("u_c", prob.u_c, "blocks", (prob.u_c + 63) // 64)
for cols in ("1", "0"):
    for T in (4, 5, 6, 7, 8, 9, 12):
        os.environ["FEBA_CHOL_COLUMNS"] = cols
        os.environ["FEBA_DAG_TILE"] = str(T)
        try:
            with fb.Handle(prob, plan=-1) as h:
                h.set_xhat(x0)
                h.iterate()
                d = h.get_delta()
                h.iterate(); h.iterate()
            print(f"columns={cols} T={T}: delta err {np.linalg.norm(d - d_ref) / np.linalg.norm(d_ref):.2e}")
        except Exception as e:
            print(f"columns={cols} T={T}: FAILED {e}")
