"""Native CSV I/O (csrc/csv_io.cpp) vs numpy on an exp_S-shaped matrix."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ursm_b200 import csvio  # noqa: E402

L, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 5000)
a = np.random.RandomState(0).randint(0, 51, (L, N)) / 50.0
d = tempfile.mkdtemp()
p1, p2 = os.path.join(d, "native.csv"), os.path.join(d, "numpy.csv")
t = time.perf_counter(); csvio.savetxt_csv(p1, a); t_wn = time.perf_counter() - t
t = time.perf_counter(); np.savetxt(p2, a, delimiter=","); t_wp = time.perf_counter() - t
same = open(p1, "rb").read() == open(p2, "rb").read()
t = time.perf_counter(); b = csvio.loadtxt_csv(p1); t_rn = time.perf_counter() - t
t = time.perf_counter(); c = np.loadtxt(p2, delimiter=","); t_rp = time.perf_counter() - t
mb = os.path.getsize(p1) / 1e6
print("%d x %d values, %.0f MB of text, %d threads; files identical: %s, arrays identical: %s"
      % (L, N, mb, os.cpu_count(), same, np.array_equal(b, c)))
print("write: native %.2f s (%.0f MB/s)   np.savetxt %.2f s (%.0f MB/s)   x%.1f"
      % (t_wn, mb / t_wn, t_wp, mb / t_wp, t_wp / t_wn))
print("read:  native %.2f s (%.0f MB/s)   np.loadtxt %.2f s (%.0f MB/s)   x%.1f"
      % (t_rn, mb / t_rn, t_rp, mb / t_rp, t_rp / t_rn))
