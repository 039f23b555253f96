"""Why BASELINE configs[2] (Burgers UQ, KSL, dt = 1e-2) cannot be run as stated once the grid is fine: counts the adaptive Tsit5
sub-steps of the three projected flows of ONE KSL step with the CPU oracle.  python tools/cfg3_stiffness.py n sqrt(m) r dt maxiters
Output committed as profiles/r02/cfg3_stiffness.txt."""
import sys, time, numpy as np, scipy.sparse as sp
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import dlra_oracle as O
n=int(sys.argv[1]); mm=int(sys.argv[2]); r=int(sys.argv[3]); dt=float(sys.argv[4]); mi=int(sys.argv[5])
nu=0.005; l=np.pi; dx=l/n; x=(np.arange(n)+0.5)*dx
i=np.arange(n)
lap=sp.csr_matrix((np.r_[np.full(n,nu/dx**2),np.full(n,-2*nu/dx**2),np.full(n,nu/dx**2)],(np.r_[i,i,i],np.r_[(i-1)%n,i,(i+1)%n])),shape=(n,n))
grad=sp.csr_matrix((np.r_[np.full(n,-0.5/dx),np.full(n,0.5/dx)],(np.r_[i,i],np.r_[(i-1)%n,(i+1)%n])),shape=(n,n))
xi=[(a,b) for a in np.linspace(-1,1,mm) for b in np.linspace(-1,1,mm)]
ub=0.5*(np.exp(np.cos(x))-1.5)*np.sin(x+2*np.pi*0.37)
rho0=np.stack([ub+0.5*a*np.sin(2*np.pi*x)+0.5*b*np.sin(3*np.pi*x) for a,b in xi],axis=1)
X0=O.truncated_svd(rho0,r)
F=lambda rho,t: lap@rho-(grad@rho)*rho
sub=lambda: O.SubStepper(maxiters=mi)
alg=O.ProjectorSplitting(O.PrimalLieTrotter(),K_alg=sub(),S_alg=sub(),L_alg=sub())
integ=O.init(O.MatrixDEProblem(F,X0,(0.0,1.0)),alg,dt)
t0=time.time()
try:
    with np.errstate(all='ignore'):
        O.step(integ)
    print("step ok |S|=",np.linalg.norm(integ.u.S))
except O.MaxItersError as e:
    print("MaxIters:",e)
c=integ.cache
print("time %.1f"%(time.time()-t0), {k:(v.naccept,v.nreject) for k,v in vars(c).items() if isinstance(v,O.SubStepper)})
print([k for k in vars(c)])
