"""Row-by-row comparison of the STIEFEL whole solve with the NumPy oracle (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import riptrm_b200 as rb
from oracle import riptrm_oracle as O
from oracle.problems import NonnegPCAStiefelProblem, nonnegpca_generate_Z
from helpers import stiefel_start

n, p, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
Z, rs = nonnegpca_generate_Z(n, seed=1)
X0 = stiefel_start(n, p, 1)
Y0 = np.ones((n, p))
opt = rb.options.default_option()
opt.update(TRS_solver="tCG", second_order_stationarity=False, maxiter=K, tolresid=0, maxtime=1e9, inner_maxiter=1000)
out = rb.RIPTRM(opt).run_stiefel(Z, X0, Y0, eps=0.01)
o = O.OracleRIPTRM({"maxiter": K, "tolresid": 0, "inner_maxiter": 1000, "manviofun": NonnegPCAStiefelProblem.manviofun})
ref = o.run(NonnegPCAStiefelProblem(Z, X0, Y0.reshape(-1), eps=0.01, closed_form=True))
L, G = out.log, ref.log
print("rows", len(L["iteration"]), len(G["iteration"]))
for i in range(min(len(L["iteration"]), len(G["iteration"]), 60)):
    f = lambda k: (float(L[k][i]) if L[k][i] is not None else float("nan"), float(G[k][i]) if G[k][i] is not None else float("nan"))
    c, r, g, nd, ap = f("cost"), f("residual"), f("gradnorm"), f("normdx"), f("ared/pred")
    print(i, L["iteration"][i], L["inner_status"][i], G["inner_status"][i], L["dxtype"][i], L["tcg_iters"][i], G["tcg_iters"][i],
          "cost %.3e" % abs(c[0] - c[1]), "res %.3e/%.3e" % r, "grad %.3e/%.3e" % g, "ndx %.2e" % abs(nd[0] - nd[1]),
          "rho %.6f/%.6f" % ap, "compl %.3e/%.3e" % f("complviolation"), "man %.1e/%.1e" % f("manviolation"))
print("final cost", L["cost"][-1], G["cost"][-1], "x diff", np.max(np.abs(out.x - ref.x)), "y diff", np.max(np.abs(out.ineqLagmult - ref.ineqLagmult)))
