"""CPU-side checks: the C-ABI library loads and exports every symbol include/ursm_b200.h
declares (no compute without a GPU), the ctypes table covers them, the product path
refuses to run without CUDA, and the multi-rank host logic works over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "ursm_b200.h")) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ursm_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ursm_b200 import _lib, build
    build.build_library()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
        assert n in _lib.SIGNATURES, "unbound in ursm_b200/_lib.py: " + n
    assert set(_lib.SIGNATURES) <= set(names)
    assert lib.ursm_version() == 101
    assert lib.ursm_launch_count() == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ursm_b200 import _lib
    from ursm_b200.utils import simplex_proj
    with pytest.raises(_lib.UrsmError):
        simplex_proj(np.ones(4), 1e-6)
    from ursm_b200.e_step_gibbs import LogitNormalGibbs_SC
    with pytest.raises(_lib.UrsmError):
        LogitNormalGibbs_SC(np.ones((4, 2)) / 4, [-1, .01], [4, .04], np.ones((3, 4)),
                            np.zeros(3, dtype=int), None)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ursm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    txt = fh.read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_shard_bounds_partition():
    from ursm_b200 import dist
    for n in (0, 1, 7, 8, 100001):
        for w in (1, 2, 3, 8):
            blocks = [dist.shard_bounds(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %r)
import torch.distributed as td
td.init_process_group(backend="gloo")
from ursm_b200 import dist
r, w = dist.rank(), dist.world_size()
assert w == 2
lo, hi = dist.shard_bounds(11)
full = np.arange(33, dtype=float).reshape(11, 3)
got = dist.gather_rows(full[lo:hi])
assert np.array_equal(got, full)
s = dist.all_reduce_np(np.array([r + 1.0, 10.0]))
assert np.allclose(s, [3.0, 20.0])
import torch
t = torch.full((4,), float(r + 1), dtype=torch.float64)
dist.all_reduce_(t)
assert torch.allclose(t, torch.full((4,), 3.0, dtype=torch.float64))
# statistics are linear in the rows: summing per-shard partials == the whole
rs = np.random.RandomState(0)
S = rs.rand(11, 5)
part = S[lo:hi].sum(axis=0)
assert np.allclose(dist.all_reduce_np(part), S.sum(axis=0))
# a rank without rows: every rank raises the same error instead of one raising and one hanging
dist.require_rows_everywhere(3, "single cells")
try:
    dist.require_rows_everywhere(0 if r == 1 else 5, "bulk samples")
    raise SystemExit("expected ValueError on rank " + str(r))
except ValueError as e:
    assert "1 of 2 ranks" in str(e), str(e)
td.destroy_process_group()
print("rank", r, "ok")
"""


def test_gloo_world2_host_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
