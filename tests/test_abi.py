"""CPU tests of the drop-in boundary: libpn2_b200.so loads and exports every symbol
include/pn2_api.h declares; the Python wrappers keep the reference's names/signatures, validate
arguments like the reference OpKernels and refuse to run without CUDA (no fallback)."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

import pointnet2_b200
from pointnet2_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pn2_api.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pn2_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pn2_api.h but not exported"
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS), "ctypes signature table and header disagree"
    assert lib.pn2_api_version() == 2
    assert os.path.dirname(_lib.lib_path()) == os.path.join(ROOT, "pointnet2_b200")  # in-tree


def test_library_is_sm100a_only():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_host_only_entry_points_work_without_gpu():
    lib = _lib.load()
    assert abs(lib.pn2_ball_threshold(0.1) - 0.01) < 1e-8
    assert lib.pn2_ball_threshold(1e-21) < 0
    assert lib.pn2_sa_layer_workspace_bytes(32, 4096, 1024, 32) >= 4 * (32 * 4096 * 3 + 32 * 1024 * (3 + 32 + 1 + 96))
    assert lib.pn2_sa_layer_workspace_bytes(0, 1, 1, 1) == 0
    assert isinstance(_lib.launch_count(), int)
    assert lib.pn2_error_string(1)  # cudaErrorInvalidValue has a message


def test_fps_plan_and_chain_override_codes_without_gpu():
    """The single-CTA plans need no device; an override names the per-step chain (pn2_api.h: cluster -1 / -2, or
    the two low bits of `threads` for cluster plans) without changing the plan that is reported."""
    lib = _lib.load()

    def plan(b, n):
        t, p, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert lib.pn2_fps_plan(b, n, ctypes.byref(t), ctypes.byref(p), ctypes.byref(c)) == 0
        return t.value, p.value, c.value

    try:
        lib.pn2_set_fps_config(0, 0, 0)
        assert plan(32, 4096) == (256, 16, 1)    # cfg2: 8 warps x 16 points
        assert plan(16, 8192) == (256, 32, 1)    # cfg4 SA1
        assert plan(32, 1024) == (128, 8, 1)     # cfg3 layer 1
        for t, p, c in ((256, 16, 1), (128, 32, 1), (512, 8, 1)):
            assert t * p * c >= 4096
        for chain in (-1, -2):
            lib.pn2_set_fps_config(256, 16, chain)
            assert plan(32, 4096) == (256, 16, 1)
            assert plan(1, 100000) == (256, 16, 1)  # an override is taken as given, whatever n
        for bits in (1, 2):
            lib.pn2_set_fps_config(128 + bits, 32, 16)
            assert plan(8, 65536) == (128, 32, 16)
        assert lib.pn2_fps_plan(0, 4096, None, None, None) != 0
    finally:
        lib.pn2_set_fps_config(0, 0, 0)


def test_argument_errors_are_reported_not_launched():
    lib = _lib.load()
    before = _lib.launch_count()
    null = ctypes.c_void_p(0)
    assert lib.pn2_fps(1, 0, 4, null, null, null, null) == 1          # n <= 0
    assert lib.pn2_fps(1, 8, 4, null, null, null, null) == 1          # null tensors
    assert lib.pn2_fps(0, 8, 4, null, null, null, null) == 0          # empty batch is a no-op
    assert lib.pn2_query_ball_point(1, 8, 2, -1.0, 4, null, null, null, null, null) == 1
    assert lib.pn2_query_ball_point(1, 8, 2, 0.1, 0, null, null, null, null, null) == 1
    assert lib.pn2_group_point(1, 0, 3, 2, 2, null, null, null, null) == 1
    assert lib.pn2_three_interpolate(1, 0, 3, 2, null, null, null, null, null) == 1
    assert _lib.launch_count() == before


REFERENCE_SIGNATURES = {
    # name: parameter names in the reference (tf_sampling.py:29,48; tf_grouping.py:8,22,33,48; tf_interpolate.py:8,19)
    "farthest_point_sample": ["npoint", "inp"],
    "gather_point": ["inp", "idx"],
    "query_ball_point": ["radius", "nsample", "xyz1", "xyz2"],
    "select_top_k": ["k", "dist"],
    "group_point": ["points", "idx"],
    "knn_point": ["k", "xyz1", "xyz2"],
    "three_nn": ["xyz1", "xyz2"],
    "three_interpolate": ["points", "idx", "weight"],
}


def test_python_surface_matches_reference_signatures():
    from pointnet2_b200 import pointnet_util, tf_grouping, tf_interpolate, tf_sampling
    mods = {"farthest_point_sample": tf_sampling, "gather_point": tf_sampling, "query_ball_point": tf_grouping,
            "select_top_k": tf_grouping, "group_point": tf_grouping, "knn_point": tf_grouping,
            "three_nn": tf_interpolate, "three_interpolate": tf_interpolate}
    for name, params in REFERENCE_SIGNATURES.items():
        fn = getattr(mods[name], name)
        assert list(inspect.signature(fn).parameters)[:len(params)] == params, name
    sg = list(inspect.signature(pointnet_util.sample_and_group).parameters)
    assert sg[:7] == ["npoint", "radius", "nsample", "xyz", "points", "knn", "use_xyz"]  # pointnet_util.py:22
    assert list(inspect.signature(pointnet_util.sample_and_group_all).parameters) == ["xyz", "points", "use_xyz"]
    assert list(inspect.signature(pointnet_util.pointnet_fp_module).parameters)[:5] == ["xyz1", "xyz2", "points1", "points2", "mlp"]
    # the module functions keep the reference's full positional order (utils/pointnet_util.py:87,156,199)
    assert list(inspect.signature(pointnet_util.pointnet_sa_module).parameters)[:16] == [
        "xyz", "points", "npoint", "radius", "nsample", "mlp", "mlp2", "group_all", "is_training", "bn_decay", "scope", "bn", "pooling",
        "knn", "use_xyz", "use_nchw"]
    assert list(inspect.signature(pointnet_util.pointnet_sa_module_msg).parameters)[:12] == [
        "xyz", "points", "npoint", "radius_list", "nsample_list", "mlp_list", "is_training", "bn_decay", "scope", "bn", "use_xyz", "use_nchw"]
    assert list(inspect.signature(pointnet_util.pointnet_fp_module).parameters)[:9] == [
        "xyz1", "xyz2", "points1", "points2", "mlp", "is_training", "bn_decay", "scope", "bn"]


def test_ops_refuse_cpu_tensors_no_fallback():
    x = torch.zeros(1, 8, 3)
    i2 = torch.zeros(1, 4, dtype=torch.int32)
    i3 = torch.zeros(1, 4, 3, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pointnet2_b200.farthest_point_sample(4, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pointnet2_b200.gather_point(x, i2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pointnet2_b200.query_ball_point(0.1, 4, x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pointnet2_b200.group_point(x, i3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pointnet2_b200.three_nn(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pointnet2_b200.three_interpolate(x, i3, torch.zeros(1, 4, 3))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            pointnet2_b200.SetAbstractionHost(1, 8, 4, 0.1, 4)


def test_attribute_and_dtype_validation_mirrors_reference():
    x = torch.zeros(1, 8, 3)
    with pytest.raises(ValueError, match="positive npoint"):      # tf_sampling.cpp:99
        pointnet2_b200.farthest_point_sample(0, x)
    with pytest.raises(ValueError, match="positive radius"):      # tf_grouping.cpp:71
        pointnet2_b200.query_ball_point(0.0, 4, x, x)
    with pytest.raises(ValueError, match="positive nsample"):     # tf_grouping.cpp:74
        pointnet2_b200.query_ball_point(0.1, 0, x, x)
    with pytest.raises(ValueError, match="positive k"):           # tf_grouping.cpp:113
        pointnet2_b200.select_top_k(0, torch.zeros(1, 2, 3))
    with pytest.raises(TypeError):
        pointnet2_b200.farthest_point_sample(4, x.double())
    with pytest.raises(TypeError):
        pointnet2_b200.farthest_point_sample(4, np.zeros((1, 8, 3), np.float32))


@pytest.mark.gpu
def test_shape_validation_on_device():
    d = torch.device("cuda:0")
    x = torch.zeros(2, 8, 3, device=d)
    with pytest.raises(ValueError, match=r"\(batch_size,num_points,3\)"):   # tf_sampling.cpp:105
        pointnet2_b200.farthest_point_sample(4, torch.zeros(2, 8, 4, device=d))
    with pytest.raises(ValueError, match="idx shape"):                      # tf_sampling.cpp:135
        pointnet2_b200.gather_point(x, torch.zeros(3, 4, dtype=torch.int32, device=d))
    with pytest.raises(ValueError, match="xyz2 shape"):                     # tf_grouping.cpp:84
        pointnet2_b200.query_ball_point(0.1, 4, x, torch.zeros(3, 4, 3, device=d))
    with pytest.raises(ValueError, match="idx shape"):                      # tf_grouping.cpp:96
        pointnet2_b200.group_point(x, torch.zeros(2, 4, dtype=torch.int32, device=d))
    with pytest.raises(ValueError, match="weight shape"):                   # tf_interpolate.cpp:206
        pointnet2_b200.three_interpolate(x, torch.zeros(2, 4, 3, dtype=torch.int32, device=d), torch.zeros(2, 5, 3, device=d))
    with pytest.raises(TypeError):
        pointnet2_b200.gather_point(x, torch.zeros(2, 4, dtype=torch.int64, device=d))
