"""ursm_b200 -- B200-native Monte-Carlo EM hot path of URSM.

Public surface mirrors the reference modules (same class / method / attribute
names):

    ursm_b200.e_step_gibbs.LogitNormalGibbs_SC / LogitNormalGibbs_BK
    ursm_b200.m_step.LogitNormalMLE
    ursm_b200.gem.LogitNormalGEM
    ursm_b200.utils.simplex_proj / std_row
    ursm_b200.scUnif  (CLI: python -m ursm_b200.scUnif  or  python scUnif.py)

All compute runs in hand-written sm_100a kernels behind the C ABI declared in
include/ursm_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
