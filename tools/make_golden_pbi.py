"""Golden fixtures of the projection-based interpolation from the CPU oracle: tests/golden/pbi_*.npz.
Each fixture holds the descriptors of two elements, the points the oracle's callback was asked for (in the order hpvert / hpedge /
hpface_opt / hpmdle_opt resp. dhpedgeE / dhpfaceE_opt / dhpfaceV_opt visit them), the function data there and the resulting dofs.
Run:  python tools/make_golden_pbi.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.test_pbi import curved_E, curved_V, gmp, random_brick, random_prism, warped_vertices  # noqa: E402

O.set_maxp(9)
O.use_blas(False)


for name, et, gen, seed in (("brick", O.MDLB, random_brick, 71), ("prism", O.MDLP, random_prism, 72)):
    rng = np.random.default_rng(seed)
    nv = 8 if et == O.MDLB else 6
    save = dict(etype=et)
    for e in range(2):
        no, noe, nof = gen(rng, 1 if e else 2, 4)
        etav = warped_vertices(rng, et)
        save[f"norder{e}"] = no; save[f"norie{e}"] = noe; save[f"norif{e}"] = nof; save[f"etav{e}"] = etav
        # H1 (update_gdof): record eta, value, eta-gradient
        rec = []

        def f_h1(eta):
            v, d = gmp(eta); rec.append((eta.copy(), v, d)); return v, d
        save[f"h1_dof{e}"] = O.pbi_element(no, noe, nof, etav[:nv], f_h1, 3, integration=0, etype=et)
        save[f"h1_eta{e}"] = np.array([r[0] for r in rec[nv:]]).reshape(-1, 3)
        save[f"h1_fvert{e}"] = np.array([r[1] for r in rec[:nv]])
        save[f"h1_fgrad{e}"] = np.array([r[2].T for r in rec[nv:]]).reshape(-1, 3, 3)      # (point, eta direction, component)
        # H(curl) (dhpedgeE, dhpfaceE_opt): record eta and the two pullbacks
        rec = []

        def f_e(eta):
            E, cE, J = curved_E(eta); rec.append((eta.copy(), (E @ J).T, np.linalg.det(J) * np.linalg.solve(J, cE.T))); return E, cE, J
        save[f"e_dof{e}"] = O.pbi_hcurl_element(no, noe, nof, etav[:nv], f_e, 2, etype=et)
        save[f"e_eta{e}"] = np.array([r[0] for r in rec]).reshape(-1, 3)
        save[f"e_fval{e}"] = np.array([r[1] for r in rec]).reshape(-1, 3, 2); save[f"e_fcurl{e}"] = np.array([r[2] for r in rec]).reshape(-1, 3, 2)
        # H(div) (dhpfaceV_opt)
        rec = []

        def f_v(eta):
            V, z, J = curved_V(eta); rec.append((eta.copy(), np.linalg.det(J) * np.linalg.solve(J, V.T))); return V, z, J
        save[f"v_dof{e}"] = O.pbi_hdiv_element(no, noe, nof, etav[:nv], f_v, 2, etype=et)
        save[f"v_eta{e}"] = np.array([r[0] for r in rec]).reshape(-1, 3)
        save[f"v_fval{e}"] = np.array([r[1] for r in rec]).reshape(-1, 3, 2)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"pbi_{name}.npz"), **save)
    print(name, {k: v.shape for k, v in save.items() if k.endswith("dof0") or k.endswith("dof1")})
