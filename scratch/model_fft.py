"""numpy model of the planned MDCT kernel arithmetic, to estimate fp32 error."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle import mdct_oracle as O

def cmul(a,b,dt):
    ar,ai=a; br,bi=b
    # (ar*br - ai*bi), (ar*bi + ai*br) with fma-like rounding approximated by plain
    return ((ar*br-ai*bi).astype(dt),(ar*bi+ai*br).astype(dt))

def small_fft(re,im,dt):
    """radix-2 DIT along last axis, in dtype dt. re,im arrays [...,n]"""
    n=re.shape[-1]
    if n==1: return re,im
    er,ei=small_fft(re[...,0::2],im[...,0::2],dt)
    orr,oi=small_fft(re[...,1::2],im[...,1::2],dt)
    k=np.arange(n//2)
    wr=np.cos(2*np.pi*k/n).astype(dt); wi=(-np.sin(2*np.pi*k/n)).astype(dt)
    tr,ti=cmul((orr,oi),(wr,wi),dt)
    return (np.concatenate([(er+tr),(er-tr)],-1).astype(dt), np.concatenate([(ei+ti),(ei-ti)],-1).astype(dt))

def fft128(re,im,dt):
    # n = j + 8 r (j=0..7, r=0..15); k = k1 + 16 k2 (k1=0..15,k2=0..7)
    sh=re.shape[:-1]
    R=re.reshape(sh+(16,8)); I=im.reshape(sh+(16,8))   # [r, j]
    R=np.swapaxes(R,-1,-2); I=np.swapaxes(I,-1,-2)     # [j, r]
    Ar,Ai=small_fft(R,I,dt)                            # [j, k1]
    j=np.arange(8)[:,None]; k1=np.arange(16)[None,:]
    ang=-2*np.pi*j*k1/128
    Ar,Ai=cmul((Ar,Ai),(np.cos(ang).astype(dt),np.sin(ang).astype(dt)),dt)
    Ar=np.swapaxes(Ar,-1,-2); Ai=np.swapaxes(Ai,-1,-2) # [k1, j]
    Vr,Vi=small_fft(Ar,Ai,dt)                          # [k1, k2]
    Vr=np.swapaxes(Vr,-1,-2).reshape(sh+(128,)); Vi=np.swapaxes(Vi,-1,-2).reshape(sh+(128,)) # k = k1+16 k2 -> [k2,k1] flat
    return Vr,Vi

def dct4_256(ue,uo,dt):
    n=np.arange(128)
    ang=-np.pi*(n+0.125)/256
    twr=np.cos(ang).astype(dt); twi=np.sin(ang).astype(dt)
    vr,vi=cmul((ue,uo),(twr,twi),dt)
    Vr,Vi=fft128(vr,vi,dt)
    yr,yi=cmul((Vr,Vi),(twr,twi),dt)
    X=np.empty(ue.shape[:-1]+(256,),dt)
    X[...,2*n]=yr; X[...,255-2*n]=-yi
    return X

def mdct_model(x, w, dt, w_lo=None):
    """x [T] fp32, w window in dt (w_lo optional low part)."""
    T=x.shape[-1]; F=T//256+1
    xp=np.pad(x,(256,256)).astype(dt)
    idx=np.arange(F)[:,None]*256+np.arange(512)[None,:]
    z=(xp[idx]*w.astype(dt)).astype(dt)
    if w_lo is not None: z=(z+xp[idx]*w_lo.astype(dt)).astype(dt)
    n=np.arange(128); lo=n<64
    g=lambda i: z[:,np.clip(i,0,511)]
    ue=np.where(lo,-g(383-2*n)-g(384+2*n), g(2*n-128)-g(383-2*n)).astype(dt)
    uo=np.where(lo, g(127-2*n)-g(128+2*n), -g(128+2*n)-g(639-2*n)).astype(dt)
    return dct4_256(ue,uo,dt)

def imdct_model(X, w, dt, w_lo=None):
    F=X.shape[0]; n=np.arange(128)
    Xd=X.astype(dt)
    U=dct4_256(Xd[:,2*n],Xd[:,255-2*n],dt)
    i=np.arange(256)
    first=np.where(i<128, U[:,np.clip(128+i,0,255)], -U[:,np.clip(383-i,0,255)])
    second=np.where(i<128,-U[:,np.clip(127-i,0,255)], -U[:,np.clip(i-128,0,255)])
    wd=w.astype(dt)
    a=(first[1:]*wd[:256]).astype(dt); b=(second[:-1]*wd[256:]).astype(dt)
    if w_lo is not None:
        a=(a+first[1:]*w_lo[:256].astype(dt)).astype(dt); b=(b+second[:-1]*w_lo[256:].astype(dt)).astype(dt)
    out=((a+b).astype(dt)*dt(4/512)).astype(dt)
    return out.reshape(-1)

if __name__=='__main__':
    g=np.load('/root/repo/tests/golden/mdct_golden.npz')
    x=g['c1_x']; w32=g['kbdwin512']; w64=O.kbdwin_f64(512)
    peak=np.abs(x).max(); eps=2.0**-23
    for name,dt in (('f64',np.float64),('f32',np.float32)):
        X=mdct_model(x,w32,dt)
        print(name,'fwd vs ref: max', np.abs(X-g['c1_spec']).max()/np.abs(g['c1_spec']).max(), 'relL2', np.linalg.norm(X-g['c1_spec'])/np.linalg.norm(g['c1_spec']))
        y=imdct_model(X,w32,dt)
        e=np.abs(y-x).max()
        print(name,'roundtrip fp32 window: max err/(eps*peak)=',e/(eps*peak),' relL2/eps=',np.linalg.norm(y-x)/np.linalg.norm(x)/eps)
        whi=w64.astype(np.float32); wlo=(w64-whi).astype(np.float32)
        X2=mdct_model(x,whi,dt,wlo); y2=imdct_model(X2,whi,dt,wlo)
        print(name,'roundtrip hi/lo window: max err/(eps*peak)=',np.abs(y2-x).max()/(eps*peak),' relL2/eps=',np.linalg.norm(y2-x)/np.linalg.norm(x)/eps)
        print(name,'fwd hi/lo vs ref relL2', np.linalg.norm(X2-g['c1_spec'])/np.linalg.norm(g['c1_spec']))
