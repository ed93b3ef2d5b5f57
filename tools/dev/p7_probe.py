"""dev probe: p=6,7 brick/prism GPU vs oracle errors and cond(G)"""
import sys, time, json, numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
from hp3d_b200.api import ElemEngine
from tests.util import hexa_xnod, uniform_order
from tests.test_gpu_prism import prism_xnod
from tests.test_oracle_prism import prism_signature
O.set_maxp(8); O.use_blas(True, threads=16)
rng = np.random.default_rng(99)
om = 2 * np.pi
prm = O.default_params(omega=om)
eng = ElemEngine(4, omega=om, maxp=8)
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
out = []
for p in (5, 6, 7):
    nob = uniform_order(p); nop = O.uniform_order(p, 3, p)
    neb = rng.integers(0, 2, 12).astype(np.int32); nfb = rng.integers(0, 8, 6).astype(np.int32)
    _, nep, nfp = prism_signature(rng, p, p)
    nHb, nHp = O.celndof(nob, 1)[0], O.celndof(nop, 3)[0]
    X = np.zeros((2, max(nHb, nHp), 3))
    X[0, :nHb] = hexa_xnod(nHb, h=0.3, jitter=0.1, rng=rng)
    X[1, :nHp] = prism_xnod(nHp, rng, h=0.3)
    res = eng.elem_stc_batch(np.stack([nob, nop]), np.stack([neb, nep]), np.stack([nfb, nfp]), X, etype=np.array([1, 3], np.int32))
    for e, (et, no, ne, nf, nH) in enumerate(((1, nob, neb, nfb, nHb), (3, nop, nep, nfp, nHp))):
        Aii, Bi, AS, BS = eng.unpack(res, e)
        t = time.time()
        rA, rB, rAS, rBS = O.condensed(4, no, ne, nf, X[e, :nH], prm, etype=et)
        t1 = time.time() - t
        O.use_blas(True, threads=1)
        rA1, rB1, _, _ = O.condensed(4, no, ne, nf, X[e, :nH], prm, etype=et)
        O.use_blas(True, threads=16)
        _, _, G, S = O.elem(4, no, ne, nf, X[e, :nH], prm, want_dpg=True, etype=et)
        Gu = np.triu(G); Gf = Gu + np.triu(Gu, 1).conj().T
        w = np.linalg.eigvalsh(Gf)
        # residual of the Schur factors
        Afull, bfull = O.elem(4, no, ne, nf, X[e, :nH], prm, etype=et)
        perm, ni, nb = O.stc_partition(4, no, et)
        Ap = Afull[np.ix_(perm, perm)]; bp = bfull[perm]
        r = dict(p=p, et=et, info=int(res["info"][e]), errA=rel(Aii, rA), errB=rel(Bi, rB), spreadA=rel(rA1, rA), spreadB=rel(rB1, rB),
                 condG=float(w[-1] / w[0]), resAS=rel(Ap[ni:, ni:] @ AS, Ap[ni:, :ni]), resBS=rel(Ap[ni:, ni:] @ BS, bp[ni:]), errAS=rel(AS, rAS),
                 condAbb=float(np.linalg.cond(Ap[ni:, ni:])), t_oracle=t1)
        print(json.dumps(r), flush=True); out.append(r)
json.dump(out, open('gpurun_out/p7_probe.json', 'w'), indent=1)
