import numpy as np
from numpy.polynomial import chebyshev as C
f32=np.float32
def fma(a,b,c): return (a.astype(np.float64)*b.astype(np.float64)+c.astype(np.float64)).astype(f32)
hi=np.pi/2*1.0001
def fit(deg):
    # fit g(u)=sin(sqrt(u))/sqrt(u) on u in [0,hi^2], weighted for relative error of sin -> minimize |x*(P-g)|/sin x = |P-g|/g
    k=np.arange(4000)+0.5
    u=(np.cos(np.pi*k/4000)+1)/2*hi*hi
    x=np.sqrt(u); g=np.sin(x)/x
    A=np.vander(u,deg+1,increasing=True)/g[:,None]
    # iteratively reweighted to approximate minimax (Lawson)
    w=np.ones_like(u)
    for it in range(200):
        c,*_=np.linalg.lstsq(A*w[:,None],np.ones_like(u)*w,rcond=None)
        r=np.abs(A@c-1)
        w=w*(r/r.max())**0.5+1e-12
        w/=w.max()
    return c
def evalf32(c,x):
    x=x.astype(f32); u=x*x
    p=np.full_like(x,f32(c[-1]))
    for ci in c[-2:0:-1]: p=fma(p,u,np.full_like(x,f32(ci)))
    # last: sin = x + x*u*p_rest  (c0 == 1 forced?) use x*(c0 + u*p)
    p=fma(p,u,np.full_like(x,f32(c[0])))
    return x*p
def evalf32_b(c,x):
    # form: x + (x*u)*q(u) with c0 forced to 1
    x=x.astype(f32); u=x*x
    q=np.full_like(x,f32(c[-1]))
    for ci in c[-2:0:-1]: q=fma(q,u,np.full_like(x,f32(ci)))
    return fma(x*u,q,x)
xs=np.linspace(1e-6,np.pi/2,2_000_001)
xs32=xs.astype(f32); tru=np.sin(xs32.astype(np.float64))
ulp=np.spacing(np.abs(tru).astype(f32)).astype(np.float64)
for deg in (4,5):
    c=fit(deg)
    print('deg',deg,'coef',[float(f32(v)) for v in c])
    for name,ev in (('a',evalf32),('b',evalf32_b)):
        y=ev(c,xs32).astype(np.float64)
        err=np.abs(y-tru)
        print('  form',name,'max rel',(err/tru).max(),'max ulp',(err/ulp).max(), 'frac correctly rounded', (y==tru.astype(f32).astype(np.float64)).mean())
# compare against glibc-ish correctly rounded: fraction equal to round(sin)
