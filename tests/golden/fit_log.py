"""How the coefficients of log_normal() (semantic-meshes_b200/csrc/smesh_fuse.cu) were fitted and checked: Q(f) of
log1p(f) = f - f^2/2 + f^3 Q(f) over f in [-1/3, 1/3] by reweighted least squares, then the float32 + FMA evaluation is
compared with log1p in double for EVERY float32 mantissa value of [2/3, 4/3) (numpy only; prints the maximum ulp error per degree)."""
import numpy as np
np.seterr(all='ignore')
# all float32 m in [2/3, 4/3)
lo=np.float32(2/3).view(np.int32); hi=np.float32(4/3).view(np.int32)
bits=np.arange(int(lo),int(hi),dtype=np.int64).astype(np.int32)
m=bits.view(np.float32)
f32=np.float32
def fma(a,b,c): return (a.astype(np.float64)*b.astype(np.float64)+c.astype(np.float64)).astype(np.float32)
f=(m-f32(1.0)).astype(np.float32)   # exact
fd=f.astype(np.float64)
true=np.log1p(fd)
# fit Q: (log1p(f) - f + f^2/2)/f^3
x=np.cos(np.pi*(np.arange(4000)+0.5)/4000)/3.0
def q_true(x): 
    return (np.log1p(x)-x+0.5*x*x)/x**3
for deg in (5,6,7):
    # minimax-ish: iteratively reweighted LS
    w=np.ones_like(x)
    for it in range(30):
        c=np.polynomial.polynomial.polyfit(x,q_true(x),deg,w=w)
        err=np.abs(np.polynomial.polynomial.polyval(x,c)-q_true(x))*np.abs(x)**3/np.maximum(np.abs(np.log1p(x)),1e-300)
        w=w*(1+err/err.max())
    cf=[f32(v) for v in c]
    # evaluate in f32 with FMA: s=f*f ; q = horner ; r = fma(q*f ... )
    q=np.full_like(f,cf[-1])
    for k in range(deg-1,-1,-1):
        q=fma(q,f,np.full_like(f,cf[k]))
    s=(f*f).astype(np.float32)          # f*f rounded
    t=fma(q,f,np.full_like(f,f32(-0.5)))  # q*f - 0.5
    r=fma(t,s,f)                          # (q f - 0.5) f^2 + f
    ulp=np.abs(r.astype(np.float64)-true)/np.spacing(np.abs(true).astype(np.float32)).astype(np.float64)
    ulp=np.where(true==0,0,ulp)
    print(deg,"max ulp",ulp.max(),"at m=",m[ulp.argmax()],"coeffs",[float(v) for v in cf])
