"""Diagnostic (not product, CPU only): d ELBO / d m at the configs[2] point (2^24 tracks, 512 x 512) evaluated with long-double
tridiagonal solves, with the dense Cholesky inverse the oracle uses and with the twisted factorisation + semiseparable recurrences
the library uses (both float64).  Printed on 2026-10-18: alpha agrees to 5e-10 / 3e-10, dm to 3e-4 / 8e-5 of the long-double
value -- the gradient is conditioning-limited, not kernel-limited."""
import numpy as np, torch, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
from oracle import vggp_oracle as O
torch.set_num_threads(8)
n=512; N=1<<24
meshes=[torch.linspace(0,1,n) for _ in range(2)]
xs,y=bench.make_tracks(0,N,N,torch.device('cpu'),torch.float32)
theta,m,Ls=bench.make_params(meshes,'cpu')
l,s2,noise=theta[:2],theta[2:4],theta[4]
K=O.kuu_b1(meshes[0],l[0],s2[0],ref_quirks=False).double()
a=torch.diagonal(K).numpy().copy(); b=torch.diagonal(K,1).numpy().copy()
def thomas(B,dt):
    aa=a.astype(dt); bb=b.astype(dt); X=B.astype(dt).copy()
    d=np.empty(n,dt); d[0]=aa[0]
    for k in range(1,n):
        w=bb[k-1]/d[k-1]; d[k]=aa[k]-w*bb[k-1]; X[k]-=w*X[k-1]
    X[n-1]/=d[n-1]
    for k in range(n-2,-1,-1): X[k]=(X[k]-bb[k]*X[k+1])/d[k]
    return X
def kron_solve(G,dt): return thomas(thomas(G,dt).T.copy(),dt).T
dd=np.empty(n); ee=np.empty(n); dd[0]=a[0]
for k in range(1,n): dd[k]=a[k]-b[k-1]**2/dd[k-1]
ee[n-1]=a[n-1]
for k in range(n-2,-1,-1): ee[k]=a[k]-b[k]**2/ee[k+1]
pd=1.0/(dd+ee-a); ru=np.zeros(n); rl=np.zeros(n); ru[:n-1]=-b/dd[:n-1]; rl[:n-1]=-b/ee[1:]
def ss1(X):
    s=pd[:,None]*X; Y=s.copy(); lc=np.zeros(X.shape[1])
    for i in range(n): Y[i]+=lc; lc=rl[i]*(s[i]+lc)
    u=np.zeros(X.shape[1])
    for i in range(n-1,-1,-1):
        Y[i]+=u; u=(ru[i-1]*(s[i]+u)) if i>0 else 0
    return Y
def kron_ss(G): return ss1(ss1(G).T.copy()).T
P=torch.cholesky_inverse(torch.linalg.cholesky(K))
def kron_dense(G): return (P@torch.from_numpy(G)@P.T).numpy()
Mm=m.reshape(n,n).numpy()
al_ld=kron_solve(Mm,np.longdouble); al_ss=kron_ss(Mm); al_de=kron_dense(Mm)
r=lambda A,B: np.linalg.norm((A-B).astype(np.float64))/np.linalg.norm(B.astype(np.float64))
print('alpha: ss vs ld',r(al_ss,al_ld),' dense vs ld',r(al_de,al_ld))
x1,x2,yy=xs[0].double(),xs[1].double(),y.double()
t=meshes[0].double(); h=(t[1:]-t[:-1])
def sten(x):
    c=(torch.searchsorted(t,x,right=False)-1).clamp(0,n-2); return c,(x-t[c])/h[c]
c1,a1=sten(x1); c2,a2=sten(x2)
def gfun(alpha):
    al=torch.from_numpy(np.asarray(alpha,dtype=np.float64))
    mu=al[c1,c2]*(1-a1)*(1-a2)+al[c1,c2+1]*(1-a1)*a2+al[c1+1,c2]*a1*(1-a2)+al[c1+1,c2+1]*a1*a2
    rr=yy-mu; g=torch.zeros(n*n,dtype=torch.float64)
    for w,di,dj in (((1-a1)*(1-a2),0,0),((1-a1)*a2,0,1),(a1*(1-a2),1,0),(a1*a2,1,1)):
        g.index_add_(0,(c1+di)*n+(c2+dj),rr*w)
    return g.reshape(n,n).numpy()
c=1.0/noise.item()
dm_ld=c*kron_solve(gfun(al_ld),np.longdouble)-al_ld
dm_ss=c*kron_ss(gfun(al_ss))-al_ss
dm_de=c*kron_dense(gfun(al_de))-al_de
print('dm: ss vs ld',r(dm_ss,dm_ld),' dense vs ld',r(dm_de,dm_ld),' ss vs dense',r(dm_ss,dm_de))
