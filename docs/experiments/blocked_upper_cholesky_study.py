"""NumPy statement of the pencil reduction in csrc/geig.cu (tnpy_geig_chol_lowest): upper Cholesky factor by row panels,
S = U^-T A U^-1 by two forward substitutions in TN form, back substitution through the diagonal blocks' inverses."""
import numpy as np
rng=np.random.default_rng(0)
n=96; B=32
Q=rng.standard_normal((n,n)); M=Q@Q.T+n*np.eye(n); A=rng.standard_normal((n,n)); A=A+A.T
G=M.copy(); nb=n//B
Cinv=[None]*nb
U=np.zeros((n,n))
for k in range(nb):
    k0=k*B
    D=G[k0:k0+B,k0:k0+B]
    L=np.linalg.cholesky(D); Ci=np.linalg.inv(L); Cinv[k]=Ci
    U[k0:k0+B,k0:k0+B]=L.T
    if k0+B<n:
        P=Ci@G[k0:k0+B,k0+B:]          # = (CinvT)^T @ G_panel  (TN)
        U[k0:k0+B,k0+B:]=P
        G[k0+B:,k0+B:]+= (-P).T@P       # TN accumulate
print("chol err",np.abs(U.T@U-M).max())
def trsm_fwd(R):   # solve U^T Y = R in place
    Y=R.copy()
    for k in range(nb):
        k0=k*B
        if k>0:
            T=U[0:k0,k0:k0+B].T@Y[0:k0,:]   # TN with K=k0
            Y[k0:k0+B,:]-=T
        Y[k0:k0+B,:]=Cinv[k]@Y[k0:k0+B,:]  # = (CinvT)^T @ .
    return Y
Y=trsm_fwd(A); Z=trsm_fwd(Y.T.copy())
S_ref=np.linalg.solve(U.T,A)@np.linalg.inv(U)
print("S err",np.abs(Z-S_ref).max()/np.abs(S_ref).max())
w,v=np.linalg.eigh(Z); z=v[:,0]
# back solve U y = z
y=np.zeros(n)
for k in range(nb-1,-1,-1):
    k0=k*B
    vv=z[k0:k0+B]-U[k0:k0+B,k0+B:]@y[k0+B:]
    y[k0:k0+B]=Cinv[k].T@vv
print("x err",np.abs(U@y-z).max(), "pencil resid", np.linalg.norm(A@y-w[0]*M@y), "norm",y@M@y)
