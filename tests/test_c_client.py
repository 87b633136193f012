"""A plain-C host of the library (tests/c_abi/abi_client.c): include/pn2_api.h must compile as C11
with gcc and link against libpn2_b200.so without torch or Python (CPU test); on a GPU box the
program runs one set-abstraction + feature-propagation pass through the C ABI on cudaMalloc'd
buffers and compares every output with the C oracle bit for bit (GPU test)."""
import os
import shutil
import subprocess

import pytest

from oracle import build as oracle_build
from pointnet2_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_client.c")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_client(tmpdir) -> str:
    lib = _build.build()
    ora = oracle_build.build_oracle()
    exe = os.path.join(str(tmpdir), "abi_client")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(CUDA, "include"), SRC, "-o", exe, lib, ora, "-L", os.path.join(CUDA, "lib64"), "-lcudart",
           f"-Wl,-rpath,{os.path.dirname(lib)}", f"-Wl,-rpath,{os.path.dirname(ora)}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not found")
def test_header_is_plain_c_and_client_links(tmp_path):
    exe = build_client(tmp_path)
    assert os.path.exists(exe)
    # the executable depends on the product library by its C symbols only
    nm = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("pn2_fps", "pn2_gather_point", "pn2_query_ball_point", "pn2_group_point", "pn2_three_nn", "pn2_three_interpolate"):
        assert f" U {sym}" in nm, sym


@pytest.mark.gpu
def test_c_client_matches_oracle_on_gpu(tmp_path):
    exe = build_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout and "DIFFERENT" not in r.stdout
