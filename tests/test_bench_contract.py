"""bench.py's reference arm and output contract, on the CPU (no GPU needed): the CPU arm
prints exactly one JSON line with the keys the driver reads, and the GPU arm refuses to run
without a device instead of falling back."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    out = _run(["--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gibbs_cell_gene_updates_per_s"
    assert d["unit"] == "updates/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("tiny")


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(["--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0"],
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        return  # on a GPU box this is the product path; covered by the gpu tests
    out = _run(["--workload", "tiny", "--steps", "1", "--warmup", "0"])
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
    assert not any(l.startswith("{") for l in out.stdout.splitlines())


def test_stdout_redirect_keeps_native_prints_off_stdout(tmp_path):
    code = (
        "import os, sys\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "with bench.StdoutToStderr():\n"
        "    os.write(1, b'native banner\\n')\n"
        "    print('python inside')\n"
        "print('{\"ok\": 1}')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == '{"ok": 1}'
    assert "native banner" in out.stderr and "python inside" in out.stderr


def test_pinned_readback_falls_back_to_plain_memory_without_cuda():
    import torch
    from ursm_b200 import _lib
    a = _lib.host_result((3, 5))
    assert a.shape == (3, 5) and a.dtype == np.float64
    a[:] = 1.0  # writable
    if not torch.cuda.is_available():
        assert a.base is None or isinstance(a.base, np.ndarray)
