"""Derivation of sincos_cw (nerf-sos_b200/csrc/common.cuh): Cody-Waite split of pi/2 into three fp32 constants and
least-squares (Chebyshev-node) fits of the sine / cosine kernels on [-pi/4, pi/4]; prints the constants and the error of an
fp32 emulation against float64 over the encoder's argument range.  tests/test_host_cpu.py re-checks the constants that are
actually in the CUDA source."""
import numpy as np
f32=np.float32
pio2=np.pi/2
C1=f32(pio2); C2=f32(pio2-np.float64(C1)); C3=f32(pio2-np.float64(C1)-np.float64(C2))
print("C", repr(float(C1)), repr(float(C2)), repr(float(C3)))
# minimax-ish fits via weighted least squares on Chebyshev nodes (double), r in [-pi/4*1.02, pi/4*1.02]
R=np.pi/4*1.01
n=4000
t=np.cos(np.pi*(np.arange(n)+0.5)/n)*R
r2=t*t
# sin(r) = r + r^3 * P(r2), P deg 3 (4 coeffs)
y=(np.sin(t)-t)/t**3
A=np.vander(r2,4,increasing=True)
ps=np.linalg.lstsq(A*(t**3/np.sin(t))[:,None]*1, y*(t**3/np.sin(t)),rcond=None)[0]   # relative-error weighting
# cos(r) = 1 + r2 * Q(r2), Q deg 3 (4 coeffs)
y2=(np.cos(t)-1)/r2
pc=np.linalg.lstsq(A*r2[:,None],y2*r2,rcond=None)[0]
ps32=ps.astype(f32); pc32=pc.astype(f32)
print("S",[repr(float(v)) for v in ps32]); print("Cc",[repr(float(v)) for v in pc32])

def fma(a,b,c):  # fp32 fma emulation through float64 (exact product, one rounding up to double-rounding rarity)
    return (a.astype(np.float64)*b.astype(np.float64)+c.astype(np.float64)).astype(f32)
def sincos_cw(x):
    x=x.astype(f32)
    q=np.rint((x*f32(2/np.pi)).astype(f32)).astype(f32)
    i=q.astype(np.int64)
    r=fma(q,np.full_like(q,-C1),x); r=fma(q,np.full_like(q,-C2),r); r=fma(q,np.full_like(q,-C3),r)
    s2=(r*r).astype(f32)
    p=np.full_like(r,ps32[3]); p=fma(p,s2,np.full_like(r,ps32[2])); p=fma(p,s2,np.full_like(r,ps32[1])); p=fma(p,s2,np.full_like(r,ps32[0]))
    t3=(s2*r).astype(f32); sn=fma(p,t3,r)
    c=np.full_like(r,pc32[3]); c=fma(c,s2,np.full_like(r,pc32[2])); c=fma(c,s2,np.full_like(r,pc32[1])); c=fma(c,s2,np.full_like(r,pc32[0]))
    cs=fma(c,s2,np.ones_like(r))
    sw=(i&1)==1
    S=np.where(sw,cs,sn); Cc=np.where(sw,sn,cs)
    S=np.where((i&2)==2,-S,S); Cc=np.where(((i+1)&2)==2,-Cc,Cc)
    return S,Cc
rng=np.random.default_rng(0)
x=(rng.uniform(-16,16,2_000_000)).astype(f32)
worst=0
for k in range(10):
    a=(x*f32(2**k)).astype(f32)
    S,Cc=sincos_cw(a)
    ts=np.sin(a.astype(np.float64)); tc=np.cos(a.astype(np.float64))
    es=np.abs(S-ts)/np.spacing(np.abs(ts).astype(f32)).astype(np.float64); ec=np.abs(Cc-tc)/np.spacing(np.abs(tc).astype(f32)).astype(np.float64)
    print(k, "max ulp err sin %.2f cos %.2f ; max abs %.2e %.2e"%(es.max(),ec.max(),np.abs(S-ts).max(),np.abs(Cc-tc).max()))
# reference: numpy float32 sin (correctly rounded-ish)
S32=np.sin((x*f32(512)).astype(f32)); 
print("np f32 sin abs err", np.abs(S32-np.sin((x*f32(512)).astype(f32).astype(np.float64))).max())
