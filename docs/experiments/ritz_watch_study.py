"""NumPy statement of csrc/ritz_watch.cuh: Sturm counts on the arrow + tridiagonal projected matrix of thick-restart
Lanczos, 256-section for the lowest eigenvalue, twisted factorisation (root = smallest |gamma|) for the eigenvector.
Run as a script: 2000 random structured matrices against numpy.linalg.eigh."""
import numpy as np
TINY=1e-290
def sturm(dg,cp,k,m,s):
    cnt=0; a=dg[k]-s
    for i in range(k):
        d=dg[i]-s
        if d==0: d=-TINY
        cnt+= d<0; a-=cp[i]*cp[i]/d
    if a==0: a=-TINY
    cnt+= a<0
    qp=1.0; q=a
    for i in range(k+1,m):
        qn=(dg[i]-s)*q-cp[i-1]**2*qp
        if qn==0: qn=-q*1e-300 if q!=0 else -TINY
        cnt+= (qn<0)!=(q<0)
        if abs(qn)>1e150: qn*=1e-150; q*=1e-150
        elif abs(qn)<1e-150: qn*=1e150; q*=1e150
        qp=q; q=qn
    return cnt
def lowest(dg,cp,k,m):
    rad=np.zeros(m)
    for i in range(m):
        if i<k: rad[i]=abs(cp[i])
        elif i==k: rad[i]=sum(abs(cp[:k]))+(abs(cp[k]) if k<m-1 else 0)
        else: rad[i]=abs(cp[i-1])+(abs(cp[i]) if i<m-1 else 0)
    lo=min(dg-rad); hi=max(dg+rad)
    w0=hi-lo; lo-=1e-3*w0+1e-300
    for rnd in range(9):
        sig=lo+(hi-lo)*(np.arange(256)+1)/257.0
        flags=[sturm(dg,cp,k,m,s)>=1 for s in sig]
        f=flags.index(True) if any(flags) else 256
        nhi= sig[f] if f<256 else hi
        nlo= sig[f-1] if f>0 else lo
        lo,hi=nlo,nhi
        if hi-lo<=1e-13*max(abs(lo),abs(hi)): break
    return lo,hi
def vec(dg,cp,k,m,s,floor):
    g=lambda x: x if x>floor else floor
    a=dg-s
    dsp=np.array([g(a[i]) for i in range(k)])          # spokes as leaves
    dm=np.zeros(m+1)                                    # bottom-up pivots of the tail, dm[j] for j>k
    for j in range(m-1,k,-1):
        v=a[j]-(cp[j]**2/dm[j+1] if j<m-1 else 0.0)
        dm[j]=g(v)
    spoke_sum=sum(cp[i]**2/dsp[i] for i in range(k))
    tail_term=(cp[k]**2/dm[k+1]) if k<m-1 else 0.0
    gam=np.zeros(m)
    gam[k]=a[k]-spoke_sum-tail_term
    dk_i=np.zeros(k)
    for i in range(k):
        dk_i[i]=g(gam[k]+cp[i]**2/dsp[i])
        gam[i]=a[i]-cp[i]**2/dk_i[i]
    dp=np.zeros(m)                                      # top-down pivots from the hub along the tail
    dp[k]=g(a[k]-spoke_sum)
    for j in range(k+1,m):
        gam[j]=a[j]-cp[j-1]**2/dp[j-1]-(cp[j]**2/dm[j+1] if j<m-1 else 0.0)
        dp[j]=g(a[j]-cp[j-1]**2/dp[j-1])
    root=int(np.argmin(np.abs(gam)))
    z=np.zeros(m)
    z[root]=1.0
    def down_from(t):
        for j in range(t,m-1): z[j+1]=-cp[j]*z[j]/dm[j+1]
    if root==k:
        for i in range(k): z[i]=-cp[i]/dsp[i]
        down_from(k)
    elif root<k:
        z[k]=-cp[root]/dk_i[root]
        for i in range(k):
            if i!=root: z[i]=-cp[i]*z[k]/dsp[i]
        down_from(k)
    else:
        for j in range(root,k,-1): z[j-1]=-cp[j-1]*z[j]/dp[j-1]
        for i in range(k): z[i]=-cp[i]*z[k]/dsp[i]
        down_from(root)
    return z/np.linalg.norm(z), root
def estimate(T,m,k,beta):
    dg=np.array([T[i,i] for i in range(m)]); cp=np.zeros(m)
    for i in range(m-1): cp[i]= T[k,i] if i<k else T[i+1,i]
    lo,hi=lowest(dg,cp,k,m)
    scale=max(abs(lo),abs(hi),np.abs(dg).max(),1e-300)
    z,root=vec(dg,cp,k,m,lo,1e-18*scale)
    return 0.5*(lo+hi), abs(beta*z[m-1]), z, root

if __name__=="__main__":
    rng=np.random.default_rng(1)
    worst=0; bad=0
    for trial in range(2000):
        m=int(rng.integers(1,49)); k=int(rng.integers(0,min(m,13))) if rng.random()<0.5 else 0
        T=np.zeros((m,m))
        scale=10**rng.uniform(-3,3)
        for i in range(m): T[i,i]=rng.standard_normal()*scale
        if k>0: T[:k,:k]=np.diag(np.sort(np.diag(T)[:k]))
        conv=10**rng.uniform(-12,0)
        for i in range(k): T[k,i]=T[i,k]=rng.standard_normal()*scale*(conv if i==0 else 10**rng.uniform(-6,0))
        for i in range(k,m-1): T[i+1,i]=T[i,i+1]=rng.standard_normal()*scale*10**rng.uniform(-6,0)
        w,v=np.linalg.eigh(T)
        th,res,z,root=estimate(T,m,k,1.0)
        ref=abs(v[m-1,0]); A=abs(w).max()
        err_th=abs(th-w[0])/A
        worst=max(worst,err_th)
        gap=(w[1]-w[0])/A if m>1 else 1
        # residual of the computed vector
        r=np.linalg.norm(T@z-th*z)/A
        if err_th>1e-13 or r>1e-12/ max(gap,1e-14) and abs(res-ref)>1e-3*ref+1e-14:
            bad+=1
            if bad<15: print("trial",trial,"m",m,"k",k,"root",root,"err_th %.1e r %.1e res %.3e ref %.3e gap %.1e"%(err_th,r,res,ref,gap))
    print("worst theta err",worst,"bad",bad)
