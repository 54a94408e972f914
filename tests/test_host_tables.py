"""Host logic of the product on the CPU: the 1-D tables the CUDA kernels are parameterised with (csrc/tables.hpp) against the
oracle's independently computed tables and against their defining identities."""
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import basis_tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dumper(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("tables") / "tables_dump")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "tables_dump.cpp"), "-o", exe])
    return exe


def load(exe, degree):
    out = subprocess.run([exe, str(degree)], capture_output=True, text=True, check=True).stdout
    t = {}
    for line in out.strip().split("\n"):
        k, *v = line.split()
        t[k] = np.array([float(x) for x in v])
    n = int(t["n"][0])
    for k in ("S", "D", "Dq", "M", "K", "Minv"):
        t[k] = t[k].reshape(n, n)
    return n, t


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7])
def test_tables_match_oracle_and_identities(dumper, degree):
    n, t = load(dumper, degree)
    o = basis_tables(degree)
    for k in ("xn", "xq", "w", "S", "D"):
        assert np.abs(t[k] - o[k]).max() < 4e-15 * max(1.0, np.abs(o[k]).max()), k  # nodes kept in long double vs double: last-bit differences
    assert np.abs(t["fd0"] - o["fd"][0]).max() < 1e-13 and np.abs(t["fd1"] - o["fd"][1]).max() < 1e-13
    one = np.ones(n)
    assert abs(t["w"].sum() - 1.0) < 1e-15
    assert np.abs(t["S"] @ one - 1.0).max() < 1e-14 and np.abs(t["D"] @ one).max() < 1e-12   # partition of unity
    assert np.abs(t["M"] - t["M"].T).max() < 1e-16 and abs(one @ t["M"] @ one - 1.0) < 1e-15      # exact mass of [0,1]
    assert np.abs(t["K"] @ one).max() < 1e-12                                                     # constants have no gradient
    assert np.abs(t["M"] @ t["Minv"] - np.eye(n)).max() < 1e-12
    # collocation basis on the Gauss points: derivative matrix differentiates polynomials exactly, traces interpolate
    p = t["xq"] ** min(degree, 3)
    dp = min(degree, 3) * t["xq"] ** (min(degree, 3) - 1)
    assert np.abs(t["Dq"] @ p - dp).max() < 1e-11
    assert abs(t["sv0"] @ p - 0.0 ** min(degree, 3)) < 1e-12 and abs(t["sv1"] @ p - 1.0) < 1e-12
    # the 1-D SIPG line operator built from these tables annihilates constants (interior faces): K 1 = 0 and jump terms vanish
    assert abs(t["fd0"] @ one) < 1e-12 and abs(t["fd1"] @ one) < 1e-12
