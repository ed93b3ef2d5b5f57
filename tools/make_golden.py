"""Generate the golden fixtures under tests/golden/ from the CPU oracle (the reference itself cannot be built here:
no Fortran compiler).  Each fixture holds the inputs the reference's `elem` would see for one or two
elements and the condensed outputs.  Run:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.util import hexa_xnod, random_signature, uniform_order  # noqa: E402

O.set_maxp(6)
O.use_blas(False)   # built-in loops: no dependence on the BLAS build
CASES = [(1, 2, 11), (1, 3, 12), (2, 2, 21), (2, 3, 22), (3, 2, 31), (4, 1, 41), (4, 2, 42)]
for kind, p, seed in CASES:
    rng = np.random.default_rng(seed)
    nel = 2
    norder = np.tile(uniform_order(p), (nel, 1))
    norie = rng.integers(0, 2, (nel, 12)).astype(np.int32)
    norif = rng.integers(0, 8, (nel, 6)).astype(np.int32)
    nH = O.celndof(norder[0])[0]
    xnod = np.stack([hexa_xnod(nH, h=0.5, origin=(0.2, 0.1 * e, 0.3), jitter=0.15, curved=0.01, rng=rng) for e in range(nel)])
    omega = 2 * np.pi if kind == 4 else (np.pi if kind == 3 else 1.0)
    prm = O.default_params(omega=omega)
    out = [O.condensed(kind, norder[e], norie[e], norif[e], xnod[e], prm) for e in range(nel)]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"cond_kind{kind}_p{p}.npz"), kind=kind, p=p, omega=omega, norder=norder,
                        norie=norie, norif=norif, xnod=xnod, Aii=np.array([o[0] for o in out]), Bi=np.array([o[1] for o in out]),
                        ASchur=np.array([o[2] for o in out]), BSchur=np.array([o[3] for o in out]))
    print(kind, p, out[0][0].shape, out[0][2].shape)

# ---- triangular prisms (element type MDLP = 3): two elements per fixture, the second with non-uniform node orders
from tests.test_gpu_prism import prism_xnod  # noqa: E402
from tests.test_oracle_prism import prism_signature  # noqa: E402

O.set_maxp(8)
for kind, p, pz, seed in [(1, 3, 2, 51), (2, 2, 2, 52), (3, 2, 3, 53), (4, 2, 2, 54)]:
    rng = np.random.default_rng(seed)
    sig = [prism_signature(rng, p, pz, uniform=(e == 0)) for e in range(2)]
    norder = np.stack([s[0] for s in sig]); norie = np.stack([s[1] for s in sig]); norif = np.stack([s[2] for s in sig])
    nHs = [O.celndof(norder[e], O.MDLP)[0] for e in range(2)]
    xnod = np.zeros((2, max(nHs), 3))
    for e in range(2):
        xnod[e, :nHs[e]] = prism_xnod(nHs[e], rng, curved=0.01)
    omega = 2 * np.pi if kind == 4 else (np.pi if kind == 3 else 1.0)
    prm = O.default_params(omega=omega)
    out = [O.condensed(kind, norder[e], norie[e], norif[e], xnod[e, :nHs[e]], prm, etype=O.MDLP) for e in range(2)]
    obj = lambda k: np.array([o[k] for o in out], dtype=object)   # noqa: E731  (elements differ in size)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"prism_kind{kind}_p{p}{pz}.npz"), kind=kind, etype=3, omega=omega,
                        norder=norder, norie=norie, norif=norif, xnod=xnod, nrdofH=np.array(nHs),
                        **{f"{n}{e}": out[e][k] for e in range(2) for k, n in enumerate(("Aii", "Bi", "ASchur", "BSchur"))})
    print("prism", kind, p, pz, out[0][0].shape, out[1][0].shape)

# ---- celem_systemI after the element: constraints, Dirichlet lift, compression on the golden elements above
from hp3d_b200 import api  # noqa: E402  (host-only entry points: hp3d_gpu_celem_pack)
from tests import celem_util as CU  # noqa: E402

O.set_maxp(6)
for kind, p, seed in [(1, 2, 61), (2, 2, 62), (3, 2, 63), (4, 1, 64)]:
    g = np.load(os.path.join(ROOT, "tests", "golden", f"cond_kind{kind}_p{p}.npz"))
    rng = np.random.default_rng(seed)
    cplx = kind >= 3
    save = dict(kind=kind, p=p)
    for e in range(2):
        c = CU.random_constraints(rng, O, api, kind, g["norder"][e], O.MDLB, cplx, dof0=1 + 500 * e)
        for k in ("nrdofl", "nrdofm_f", "idbc", "zdofd", "nextract", "lcon", "cptr", "cidx", "cval"):
            save[f"{k}{e}"] = np.asarray(c[k])
        for f in range(3):
            save[f"nrcon{e}_{f}"] = c["nrcon"][f]; save[f"nac{e}_{f}"] = c["nac"][f]; save[f"constr{e}_{f}"] = c["constr"][f]
        for isym in (1, 2, 3):
            zb, za = CU.oracle_celem(O, c, g["Aii"][e], g["Bi"][e], isym)
            save[f"zbload{e}"] = zb; save[f"zastif{e}_{isym}"] = za
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"celem_kind{kind}_p{p}.npz"), **save)
    print("celem", kind, p, save["zastif0_2"].shape)
