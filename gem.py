"""Drop-in for `from gem import *` (gem.py): re-exports the sm_100a-backed classes
under the reference's module name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ursm_b200.gem import LogitNormalGEM  # noqa: E402,F401
from ursm_b200.e_step_gibbs import LogitNormalGibbs_BK, LogitNormalGibbs_SC  # noqa: E402,F401
from ursm_b200.m_step import LogitNormalMLE  # noqa: E402,F401
from ursm_b200.utils import simplex_proj, std_row  # noqa: E402,F401
