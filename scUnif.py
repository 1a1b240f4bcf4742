#!/usr/bin/env python
"""Drop-in for the reference's `python scUnif.py ...` (scUnif.py:101-249):
same flags, same gemout_*.csv outputs; the fit runs on the sm_100a kernels of
ursm_b200.  See ursm_b200/scUnif.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ursm_b200.scUnif import main  # noqa: E402

if __name__ == "__main__":
    main()
