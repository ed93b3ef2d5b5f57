"""GPU check of the dense phase (Cholesky/solve/HERK/condensation) against numpy on oracle-produced Gram/stiffness."""
import ctypes as C, sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
from hp3d_b200 import _lib
L = _lib.lib(); _lib.check(L.hp3d_gpu_init(0))
def cube_xnod(nH, h=0.25, jitter=0.15, seed=0):
    V=np.array([[0,0,0],[1,0,0],[1,1,0],[0,1,0],[0,0,1],[1,0,1],[1,1,1],[0,1,1]],float)
    X=np.zeros((nH,3)); X[:8]=h*V+np.random.default_rng(seed).uniform(-jitter*h,jitter*h,(8,3)); return X
O.use_blas(True)
def run(p, nel=2):
    no=O.uniform_order(p); oe=np.zeros(12,int); of=np.zeros(6,int); nH=O.celndof(no)[0]
    prm=O.default_params(omega=2*np.pi)
    Gs=[];Bs=[];refs=[]
    for e in range(nel):
        X=cube_xnod(nH,seed=e)
        A,b,G,S=O.elem(O.MAXW_UW,no,oe,of,X,prm,want_dpg=True)
        perm,ni,nb=O.stc_partition(O.MAXW_UW,no)
        n=G.shape[0]
        # columns [bubble | interface | load]
        cols=list(range(ni,ni+nb))+list(range(ni))+[ni+nb]
        Gs.append(np.asfortranarray(G)); Bs.append(np.asfortranarray(S[:,cols]))
        refs.append(O.condensed(O.MAXW_UW,no,oe,of,X,prm))
    G=np.stack([g.T for g in Gs]).copy(); B=np.stack([b.T for b in Bs]).copy()  # per element column-major
    Aii=np.zeros((nel,ni,ni),complex); Bi=np.zeros((nel,ni),complex); AS=np.zeros((nel,ni,nb),complex); BS=np.zeros((nel,nb),complex)
    info=np.zeros(nel,np.int32)
    t=time.time()
    rc=L.hp3d_gpu_dense_debug(1,nel,n,nb,ni,G.ctypes.data_as(C.c_void_p),B.ctypes.data_as(C.c_void_p),Aii.ctypes.data_as(C.c_void_p),Bi.ctypes.data_as(C.c_void_p),AS.ctypes.data_as(C.c_void_p),BS.ctypes.data_as(C.c_void_p),info.ctypes.data_as(C.c_void_p))
    _lib.check(rc)
    for e in range(nel):
        r=refs[e]; got=(Aii[e].T,Bi[e],AS[e].T,BS[e])
        print(p,e,"info",info[e],"relerr",["%.2e"%(np.linalg.norm(a-b)/np.linalg.norm(a)) for a,b in zip(r,got)], "t=%.2f"%(time.time()-t))
for p in (2,3,5): run(p)
