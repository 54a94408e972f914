"""The header-only C++ shim (the reference is C++): compile a translation unit against it, link libexadg_b200.so and run it.
Without a GPU the program only exercises the error path; on the GPU box it runs vmult / diagonal through the shim."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


def _build_and_run(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    exe = str(tmp_path / "shim_smoke")
    libdir = os.path.join(ROOT, "exadg_b200")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "shim_smoke.cpp"),
           "-o", exe, "-L", libdir, "-lexadg_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir + ",-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "SHIM_OK" in out.stdout
    return out.stdout


def test_cpp_shim_compiles_links_and_runs(tmp_path):
    _build_and_run(tmp_path)


@pytest.mark.gpu
def test_cpp_shim_runs_vmult_on_the_gpu(tmp_path):
    out = _build_and_run(tmp_path)
    assert "max|A*1|" in out and "no GPU" not in out
