"""Test data for the constrained-approximation / compression step (celem_systemI.F90:543-785).

`random_constraints` fabricates what `logic` + the extraction loops of celem_systemI hand to that step for one element
(nrcon/nac/constr per family, IDBC/ZDOFD, NEXTRACT, LCON).  `HangingMesh` is a real 1-irregular mesh: a coarse N^3 grid of
p=1 bricks with one brick split into eight, so that edge midpoints / face centres shared with unrefined neighbours are
hanging nodes (value = mean of the 2 / 4 parent vertices), the situation the transform exists for.
"""
import numpy as np

from tests.util import VERT


def single_component_dofs(O, kind, norder, etype):
    """(nrdoflHi, nrdoflEi, nrdoflVi) of celem_systemI.F90:104-113 restricted to the families the problem uses."""
    norderi = np.array(norder, dtype=np.int32).copy()
    norderi[O.mid_index(etype)] = 11 if etype == O.MDLP else 111
    h, e, v, _ = O.celndof(norderi, etype)
    use = {O.POIS_GAL: (1, 0, 0), O.POIS_PDPG: (1, 0, 1), O.MAXW_GAL: (0, 1, 0), O.MAXW_UW: (0, 1, 0)}[kind]
    return [h * use[0], e * use[1], v * use[2]]


def random_constraints(rng, O, api, kind, norder, etype, cplx, nacdim=4, frac_con=0.3, frac_dbc=0.2, dof0=1, extra=5):
    """One element's constraint data: a fraction of the element dofs is constrained to 2..nacdim random modified dofs, the
    others map one-to-one; a fraction of the modified dofs is Dirichlet with random data; NEXTRACT = the rest, shuffled."""
    pho, php = O.physics_of(kind), api.physics_default(kind)
    nrdofl = single_component_dofs(O, kind, norder, etype)
    nrcon, nac, constr, nrdofm_f = [], [], [], []
    for f in range(3):
        nk = nrdofl[f]
        nm = nk + (extra if nk else 0)            # single-component modified dofs of this family
        rc = np.ones(nk, np.int32); na = np.zeros((nk, nacdim), np.int32); co = np.zeros((nk, nacdim))
        perm = rng.permutation(nm)[:nk] + 1 if nk else np.zeros(0, int)
        for k in range(nk):
            if rng.random() < frac_con:
                m = int(rng.integers(2, nacdim + 1))
                rc[k] = m
                na[k, :m] = rng.choice(nm, m, replace=False) + 1
                co[k, :m] = rng.uniform(-1, 1, m)
            else:
                na[k, 0] = perm[k]; co[k, 0] = 1.0
        nrcon.append(rc); nac.append(na); constr.append(co)
        nrdofm_f.append(nm * pho.nrvar[f])
    nrdofm = int(sum(nrdofm_f))
    idbc = (rng.random(nrdofm) < frac_dbc).astype(np.int32)
    zd = rng.standard_normal(nrdofm) + (1j * rng.standard_normal(nrdofm) if cplx else 0)
    zd = np.where(idbc == 1, zd, 0).astype(np.complex128 if cplx else np.float64)
    nextract = (rng.permutation(np.flatnonzero(idbc == 0)) + 1).astype(np.int32)
    lcon = (dof0 + rng.permutation(4 * len(nextract))[: len(nextract)]).astype(np.int32)
    cptr, cidx, cval = api.celem_pack(php, nrdofl, nrcon, nac, constr, nrdofm_f)
    return dict(nrdofl=nrdofl, nrcon=nrcon, nac=nac, constr=constr, nrdofm_f=nrdofm_f, idbc=idbc, zdofd=zd, nextract=nextract, lcon=lcon,
                cptr=cptr, cidx=cidx, cval=cval, ph=pho)


def oracle_celem(O, c, A, b, isym):
    return O.celem_modify(c["ph"], c["nrdofl"], c["nrcon"], c["nac"], c["constr"], c["nrdofm_f"], A, b, c["idbc"], c["zdofd"], c["nextract"], isym)


class HangingMesh:
    """N^3 coarse p=1 bricks on [0,1]^3; brick `ref` is split into 8 children.  Global dofs = regular vertices."""

    def __init__(self, N=2, ref=(0, 0, 0)):
        self.N = N
        h = 1.0 / N
        cells = [(i, j, k) for k in range(N) for j in range(N) for i in range(N)]
        self.elems = []      # (origin, size)
        for c in cells:
            if c == tuple(ref):
                for v in VERT:
                    self.elems.append((np.array(c) * h + v * h / 2, h / 2))
            else:
                self.elems.append((np.array(c) * h, h))
        r0 = np.array(ref) * h
        # regular vertices: the coarse grid + those lattice points of the refined brick that no unrefined brick touches
        def key(x):
            return tuple(int(round(v * 4 * N)) for v in x)
        self.key = key
        reg = {}
        for i in range(N + 1):
            for j in range(N + 1):
                for k in range(N + 1):
                    reg.setdefault(key(np.array([i, j, k]) * h), len(reg))
        self.parents = {}    # hanging vertex key -> list of (regular key, coefficient)
        for a in range(3):
            for b in range(3):
                for c_ in range(3):
                    t = np.array([a, b, c_])
                    x = r0 + t * h / 2
                    if key(x) in reg:
                        continue
                    mids = [d for d in range(3) if t[d] == 1]
                    # the point lies in the relative interior of the entity spanned by the `mids` directions; it is shared with
                    # an unrefined neighbour iff that entity is not entirely on the domain boundary / interior of the brick
                    if len(mids) == 3:
                        reg[key(x)] = len(reg)            # centre of the refined brick
                        continue
                    # the point is the midpoint of an edge of the brick (one mid direction) or the centre of one of its faces (two);
                    # an unrefined neighbour shares that edge / face unless all its fixed coordinates lie on the domain boundary
                    fixed = [d for d in range(3) if t[d] != 1]
                    shared = any(0 < x[d] < 1 for d in fixed)
                    if not shared:
                        reg[key(x)] = len(reg)
                        continue
                    par = []
                    for s in np.ndindex(*([2] * len(mids))):
                        y = x.copy()
                        for d, sd in zip(mids, s):
                            y[d] += (2 * sd - 1) * h / 2
                        par.append((key(y), 1.0 / 2 ** len(mids)))
                    self.parents[key(x)] = par
        self.reg = reg
        self.ndof = len(reg)
        self.xyz = np.zeros((self.ndof, 3))
        for kx, g in reg.items():
            self.xyz[g] = np.array(kx) / (4.0 * N)
        self.bdry = np.array([any(abs(v) < 1e-12 or abs(v - 1) < 1e-12 for v in self.xyz[g]) for g in range(self.ndof)])

    def descriptors(self):
        nel = len(self.elems)
        X = np.zeros((nel, 8, 3))
        for e, (o, s) in enumerate(self.elems):
            X[e] = o + s * VERT
        no = np.tile(np.array([1] * 12 + [11] * 6 + [111], np.int32), (nel, 1))
        return no, np.zeros((nel, 12), np.int32), np.zeros((nel, 6), np.int32), X

    def constraints(self, api, O, uex, free_numbering):
        """Per element: logic-like arrays for the H1 family + IDBC/ZDOFD/NEXTRACT/LCON; free_numbering: global dof -> 1-based
        equation number of the non-Dirichlet dofs."""
        pho, php = O.physics_of(O.POIS_GAL), api.physics_default(O.POIS_GAL)
        out = []
        for (o, s) in self.elems:
            nodm, nrcon, nac, con = [], np.zeros(8, np.int32), np.zeros((8, 4), np.int32), np.zeros((8, 4))
            def mod_index(kx):
                if kx not in nodm:
                    nodm.append(kx)
                return nodm.index(kx) + 1
            for k in range(8):
                kx = self.key(o + s * VERT[k])
                if kx in self.reg:
                    nrcon[k] = 1; nac[k, 0] = mod_index(kx); con[k, 0] = 1.0
                else:
                    par = self.parents[kx]
                    nrcon[k] = len(par)
                    for q, (pk, cf) in enumerate(par):
                        nac[k, q] = mod_index(pk); con[k, q] = cf
            nm = len(nodm)
            g = np.array([self.reg[kx] for kx in nodm])
            idbc = self.bdry[g].astype(np.int32)
            zd = np.where(idbc == 1, np.array([uex(self.xyz[q]) for q in g]), 0.0)
            nextract = (np.flatnonzero(idbc == 0)[::-1] + 1).astype(np.int32)      # reversed node order, as celem_systemI loops
            lcon = np.array([free_numbering[g[l - 1]] for l in nextract], np.int32)
            nrdofl, nrdofm_f = [8, 0, 0], [nm, 0, 0]
            z = np.zeros((0, 4))
            cptr, cidx, cval = api.celem_pack(php, nrdofl, [nrcon, [], []], [nac, z, z], [con, z, z], nrdofm_f)
            out.append(dict(nrdofl=nrdofl, nrcon=[nrcon, np.zeros(0, np.int32), np.zeros(0, np.int32)], nac=[nac, z.astype(np.int32), z.astype(np.int32)],
                            constr=[con, z, z], nrdofm_f=nrdofm_f, idbc=idbc, zdofd=zd, nextract=nextract, lcon=lcon, cptr=cptr, cidx=cidx,
                            cval=cval, ph=pho))
        return out
