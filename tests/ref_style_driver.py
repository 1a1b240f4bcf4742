"""A driver that uses ONLY the reference's API surface of the hot-path classes, in the order
and with the argument forms `gem.py` of the reference uses them (gem.py:137-149 estep_gibbs,
:152-195 gem, :198-205 init_mle, :283-295 init_gibbs) -- no `seed`, `row_offset` or device
arguments, the merged `suff_stats` dict built with `dict.update`, attributes read back as
the reference reads them.  Test infrastructure: the GPU box has no /root/reference, so this
stands in for exec'ing the reference's gem.py on top of the ursm_b200 classes there
(tests/test_dropin_reference_gem.py does exactly that where the reference tree exists)."""
import numpy as np


def fit(classes, BKexpr, SCexpr, G, K, init_A, init_alpha, init_pkappa, init_ptau, iMarkers=None,
        burnin=10, sample=20, thin=1, bk_mean_approx=True, est_alpha=True, min_A=1e-6,
        MLE_CONV=1e-6, MLE_maxiter=100, EM_CONV=1e-3, EM_maxiter=10):
    LogitNormalGibbs_SC, LogitNormalGibbs_BK, LogitNormalMLE = classes
    hasBK, hasSC = BKexpr is not None, SCexpr is not None
    itype = [np.where(G == k)[0] for k in range(K)] if hasSC else None
    A, alpha = np.copy(init_A), (np.copy(init_alpha) if hasBK else None)
    pkappa, ptau = (np.copy(init_pkappa), np.copy(init_ptau)) if hasSC else (None, None)
    # gem.py:283-295
    if hasSC:
        Gibbs_SC = LogitNormalGibbs_SC(A=init_A, pkappa=init_pkappa, ptau=init_ptau,
                                       SCexpr=SCexpr, G=G, itype=itype)
        Gibbs_SC.init_gibbs()
    if hasBK:
        Gibbs_BK = LogitNormalGibbs_BK(A=init_A, alpha=init_alpha, BKexpr=BKexpr,
                                       iMarkers=iMarkers)
        Gibbs_BK.init_gibbs()
    # gem.py:198-205
    mle = LogitNormalMLE(BKexpr=BKexpr, SCexpr=SCexpr, G=G, K=K, itype=itype, hasBK=hasBK,
                         hasSC=hasSC, init_A=init_A, init_alpha=init_alpha,
                         init_pkappa=init_pkappa, init_ptau=init_ptau, min_A=min_A,
                         MLE_CONV=MLE_CONV, MLE_maxiter=MLE_maxiter)
    converged, niter, old_elbo = EM_CONV + 1, 0, -10 ** 6
    path_elbo = np.array([])
    suff_stats = {}
    while abs(converged) > EM_CONV and niter < EM_maxiter:
        suff_stats = {}
        if hasSC:  # gem.py:139-143
            Gibbs_SC.update_parameters(A, pkappa, ptau)
            Gibbs_SC.gibbs(burnin=burnin, sample=sample, thin=thin)
            suff_stats.update(Gibbs_SC.suff_stats)
        if hasBK:  # gem.py:145-149
            Gibbs_BK.update_parameters(A, alpha)
            Gibbs_BK.gibbs(burnin=burnin, sample=sample, thin=thin, mean_approx=bk_mean_approx)
            suff_stats.update(Gibbs_BK.suff_stats)
        mle.update_suff_stats(suff_stats)  # gem.py:167
        elbo = mle.compute_elbo()
        if hasSC:
            mle.opt_kappa_tau()
            pkappa, ptau = mle.pkappa, mle.ptau
        if hasBK and est_alpha:
            mle.opt_alpha()
            alpha = mle.alpha
        mle.opt_A_u()
        A = mle.A
        elbo = mle.compute_elbo()
        converged = abs(elbo - old_elbo)
        old_elbo = elbo
        niter += 1
        path_elbo = np.append(path_elbo, [elbo])
    return dict(A=A, alpha=alpha, pkappa=pkappa, ptau=ptau, path_elbo=path_elbo,
                suff_stats=suff_stats, niter=niter)
