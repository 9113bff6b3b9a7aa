"""The estimate of ritz_watch_study.py on the T of a real thick-restart Lanczos run (ncv 32, keep 10) across restarts,
against a dense eigh of T at every step."""
import numpy as np, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from ritz_watch_study import estimate
rng=np.random.default_rng(3)
N=600
A=rng.standard_normal((N,N)); A=(A+A.T)/2; ev=np.linspace(-1,1,N)**3*50; 
Q,_=np.linalg.qr(rng.standard_normal((N,N))); A=Q@np.diag(ev)@Q.T
ncv,keep=32,10
V=np.zeros((ncv+1,N)); v=rng.standard_normal(N); V[0]=v/np.linalg.norm(v)
T=np.zeros((48,48)); j=0;k_arrow=0; worst=0; n=0
while n<400:
    w=A@V[j]
    h=V[:j+1]@w; w-=V[:j+1].T@h; h2=V[:j+1]@w; w-=V[:j+1].T@h2
    T[:j+1,j]=h+h2; T[j,:j+1]=h+h2
    beta=np.linalg.norm(w); V[j+1]=w/beta; n+=1
    m=j+1
    Tm=T[:m,:m]; wv,S=np.linalg.eigh(Tm)
    ref=abs(beta*S[m-1,0])
    th,res,z,root=estimate(T,m,k_arrow,beta)
    rel=abs(res-ref)/max(ref,1e-300)
    anorm=abs(wv).max()
    if ref>1e-13*anorm: worst=max(worst,rel)
    if rel>1e-3 and ref>1e-13*anorm: print("step",n,"m",m,"k",k_arrow,"root",root,"res %.3e ref %.3e th err %.1e"%(res,ref,abs(th-wv[0])))
    if ref<=1e-10*anorm:
        print("converged at",n,"resid",ref, "est", res); break
    if m==ncv:
        Y=S[:,:keep].T@V[:m]
        V[keep]=V[m]; V[:keep]=Y
        T[:]=0; T[np.arange(keep),np.arange(keep)]=wv[:keep]
        j=keep; k_arrow=keep
    else: j+=1
print("worst rel diff of residual estimate",worst)
