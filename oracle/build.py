"""Builds the oracle's C part (gcc, OpenMP) into oracle/_build/ and the host build of
the sampler header the CPU tests use.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.build

The reference itself is pure Python (no C/C++ source under /root/reference), so
there is nothing to compile into oracle/_ref; its one native dependency
(pypolyagamma 1.1.1) is not vendored -- pg_devroye.c restates that algorithm.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
PG_SRC = os.path.join(HERE, "pg_devroye.c")
PG_LIB = os.path.join(OUT, "libpg_devroye.so")


def _stale(target, deps):
    return (not os.path.exists(target)
            or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps))


def build_pg(force=False):
    os.makedirs(OUT, exist_ok=True)
    if force or _stale(PG_LIB, [PG_SRC]):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", PG_LIB, PG_SRC,
                               "-lm"])
    return PG_LIB


def build_all(force=False):
    return build_pg(force)


if __name__ == "__main__":
    print(build_all(force=True))
