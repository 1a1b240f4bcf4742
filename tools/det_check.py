import numpy as np, sys, hashlib
sys.path.insert(0,".")
import ursm_b200.e_step_gibbs as es, ursm_b200.m_step as ms
from oracle import ursm_oracle as orc
d=np.load("tests/golden/demo_data.npz")
Y,X,G=d["Y"],d["X"],d["G"]; K=3; N=Y.shape[1]
itype=[np.where(G==k)[0] for k in range(K)]
A=orc.init_A(Y,X,K,itype,1e-6)
pk,pt,al=np.array([-1.,.01]),np.array([N,.01*N],dtype=float),np.ones(K)
h=lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:8]
for rep in range(3):
    sc=es.LogitNormalGibbs_SC(A,pk,pt,Y,G,itype,seed=5)
    sc.gibbs(10,20,1)
    st=sc.suff_stats
    bk=es.LogitNormalGibbs_BK(A,al,X,seed=5); bk.init_gibbs(); bk.gibbs(mean_approx=True)
    allst=dict(st); allst.update(bk.suff_stats)
    mle=ms.LogitNormalMLE(X,Y,G,K,itype,True,True,A,al,pk,pt,1e-6,1,1e-6,500)
    mle.update_suff_stats(allst)
    e1=mle.compute_elbo()
    mle.opt_kappa_tau(); mle.opt_alpha(); mle.opt_A_u()
    print(rep,"expS",h(np.asarray(st["exp_S"])),"kap",h(st["exp_kappa"]),"cA",h(st["coeffA"]),"cQ",h(st["coeffAsq"]),"elboc",repr(st["exp_elbo_const"]),
          "| W",h(bk.suff_stats["exp_W"]),"Zik",h(bk.suff_stats["exp_Zik"]),"| u",h(mle.u),"elbo",repr(e1),"A",h(mle.A),"niter",mle.niter_A, repr(mle.compute_elbo()))
