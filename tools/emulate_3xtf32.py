"""Numpy emulation of the 3xTF32 split (hi = truncation or round-to-nearest to TF32, lo = x - hi, products hi*hi + lo*hi + hi*lo, FP32 accumulation)
applied to the blocked Cholesky of the row-GP kernel: error of mean / variance / alpha against float64 on the bench distribution (C4)."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle import oracle_np
def trunc_tf32(x):
    return (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
def rna_tf32(x):
    return ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)
def mm3(a,b,mode):  # a (m,k) b (k,n) float32, emulated 3xTF32, fp32 accumulate (numpy float32 matmul accumulates in fp32-ish)
    if mode=='fp32': return a@b
    f = trunc_tf32 if mode=='trunc' else rna_tf32
    ah=f(a); al=trunc_tf32((a-ah).astype(np.float32)); bh=f(b); bl=trunc_tf32((b-bh).astype(np.float32))
    return (al@bh + ah@bl + ah@bh).astype(np.float32)
def chol_blocked(K,y,mode):
    n=K.shape[0]; L=np.zeros_like(K); nb=n//16
    Dinv=[]
    for kb in range(nb):
        c0=16*kb
        S = mm3(L[c0:, :c0], L[c0:c0+16,:c0].T.copy(), mode) if kb>0 else np.zeros((n-c0,16),np.float32)
        P = (K[c0:, c0:c0+16]-S).astype(np.float32)
        D = np.linalg.cholesky(P[:16].astype(np.float64)).astype(np.float32)  # pivot block (fp32-ish)
        Di = np.linalg.inv(D.astype(np.float64)).astype(np.float32)
        Dinv.append(Di)
        L[c0:c0+16,c0:c0+16]=D
        if c0+16<n:
            L[c0+16:, c0:c0+16] = mm3(P[16:], Di.T.copy(), mode)
    return L, Dinv
def solve(L,y):
    import scipy.linalg as sl
    z=sl.solve_triangular(L.astype(np.float32),y.astype(np.float32),lower=True).astype(np.float32)
    a=sl.solve_triangular(L.T.astype(np.float32),z,lower=False).astype(np.float32)
    return a
rng=np.random.default_rng(6)
errs={m:[] for m in ('fp32','trunc','rna')}
for trial in range(40):
    n,d,t=128,3,128
    x=rng.random((n,d),dtype=np.float32); wv=rng.uniform(1,4,(1,d)).astype(np.float32)
    y=(0.5*np.sin(wv*x*3.0).sum(axis=1)).astype(np.float32)
    var=np.full(n,0.01,np.float32)
    xq=rng.random((t,d),dtype=np.float32)
    K64=oracle_np.ktrain(oracle_np.MATERN32,0.3,x.astype(np.float64),var.astype(np.float64))
    kt64=oracle_np.ktest(oracle_np.MATERN32,0.3,x.astype(np.float64),xq.astype(np.float64))
    import scipy.linalg as sl
    c=sl.cho_factor(K64,lower=True); a64=sl.cho_solve(c,y.astype(np.float64)); m64=kt64.T@a64
    v64=1-(sl.solve_triangular(c[0],kt64,lower=True)**2).sum(0)
    for mode in errs:
        L,_=chol_blocked(K64.astype(np.float32),y,mode)
        a=solve(L,y)
        m=(kt64.astype(np.float32).T@a)
        V=sl.solve_triangular(L.astype(np.float64),kt64,lower=True)
        v=1-(V**2).sum(0)
        errs[mode].append((np.abs(m-m64).max()/np.abs(m64).max(), np.abs(v-v64).max(), np.abs(a-a64).max()/np.abs(a64).max()))
for m,e in errs.items():
    e=np.array(e); print(m,'mean err max %.2e  var err max %.2e  alpha rel err max %.2e'%tuple(e.max(0)), ' medians %.2e %.2e %.2e'%tuple(np.median(e,0)))
