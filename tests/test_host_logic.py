"""CPU tier: host logic of the package, the C ABI surface, and the kernels' schedules under CPU emulation."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import recconv_oracle as O
from tests.helpers import GOLDEN, TOL_BF16, TOL_FP32, load_recconv_golden, recconv_golden_files, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from recnext_b200 import build, _native

    build.build()
    return _native


def test_library_exports_every_declared_symbol(native):
    """Every function declared in include/recnext_b200.h is exported by the built library (no compute calls)."""
    hdr = open(os.path.join(ROOT, "include", "recnext_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rec\w+)\s*\(", hdr))
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    L = native.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.recnext_abi_version() == native.ABI_VERSION


def test_source_index_helper_bit_exact_vs_torch_tables(native):
    z = np.load(os.path.join(GOLDEN, "interp_index_tables.npz"))
    L = native.lib()
    for i, o in z["pairs"]:
        i, o = int(i), int(o)
        i0 = np.zeros(o, np.int32); i1 = np.zeros(o, np.int32); lam = np.zeros(o, np.float32)
        assert L.recconv_source_index(1, i, o, i0.ctypes.data, None, None) == 0
        np.testing.assert_array_equal(i0, z[f"nearest_{i}_{o}"])
        assert L.recconv_source_index(0, i, o, i0.ctypes.data, i1.ctypes.data, lam.ctypes.data) == 0
        for d in range(o):
            a, b, l = O.bilinear_index(i, o, d)
            assert (a, b) == (int(i0[d]), int(i1[d])) and l == float(lam[d]), (i, o, d)
    # the fused-multiply-add tie case (see oracle/recconv_oracle.c): 129 -> 257, dst 128
    i0 = np.zeros(257, np.int32); i1 = np.zeros(257, np.int32); lam = np.zeros(257, np.float32)
    L.recconv_source_index(0, 129, 257, i0.ctypes.data, i1.ctypes.data, lam.ctypes.data)
    assert i0[128] == 63 and lam[128] > 0.9999


def test_bad_arguments_are_errors(native):
    L = native.lib()
    d = native.RecConvDesc(1, 4, 8, 8, 4, 1, 0, 0, 0, 0)  # even kernel
    buf = ctypes.create_string_buffer(256)
    assert L.recconv_plan_describe(ctypes.byref(d), 0, buf, 256) == -1
    assert b"kernel_size" in L.recnext_last_error()
    d = native.RecConvDesc(1, 4, 8, 8, 5, 9, 0, 0, 0, 0)  # level too deep
    assert L.recconv_plan_describe(ctypes.byref(d), 0, buf, 256) == -1
    d = native.RecConvDesc(2, 64, 200, 336, 5, 4, 0, 1, 0, 0)  # detection stage 0: pyramid exceeds shared memory -> streamed path
    assert L.recconv_plan_describe(ctypes.byref(d), 1, buf, 256) == 0 and b"streamed" in buf.value
    assert L.recconv_backward_workspace_bytes(ctypes.byref(d)) > 2 * 64 * 200 * 336 * 4
    d = native.RecConvDesc(256, 64, 56, 56, 5, 4, 0, 1, 0, 0)  # the fused kernels need no forward workspace
    assert L.recconv_forward_workspace_bytes(ctypes.byref(d)) == 0


def test_plan_describe_baseline_shapes(native):
    import recnext_b200 as R

    for shape, lvl in [((256, 64, 56, 56), 4), ((256, 128, 28, 28), 3), ((256, 256, 14, 14), 2), ((256, 512, 7, 7), 1)]:
        for bwd in (False, True):
            s = R.plan_describe(shape, 5, lvl, "bilinear", torch.bfloat16, False, bwd)
            assert "tma=1" in s, s


def test_module_state_dict_matches_reference_layout():
    import recnext_b200 as R

    m = R.RecConv2d(8, kernel_size=5, bias=True, level=3, mode="nearest")
    keys = list(m.state_dict().keys())
    assert keys == ["down.weight", "down.bias"] + [f"convs.{j}.{n}" for j in range(4) for n in ("weight", "bias")]
    assert all(tuple(v.shape) in ((8, 1, 5, 5), (8,)) for v in m.state_dict().values())
    m2 = R.RecConv2d(8, level=2)
    assert list(m2.state_dict().keys()) == ["down.weight", "convs.0.weight", "convs.1.weight", "convs.2.weight"]
    # same default init stream as the reference (nn.Conv2d in the same construction order)
    torch.manual_seed(0)
    a = R.RecConv2d(4, level=1)
    torch.manual_seed(0)
    down = torch.nn.Conv2d(4, 4, 5, 2, 2, groups=4, bias=False)
    assert torch.equal(a.down.weight, down.weight)


def test_no_cpu_fallback():
    import recnext_b200 as R

    m = R.RecConv2d(4, level=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 4, 8, 8))
    with pytest.raises(ValueError):
        R.RecConv2d(4, kernel_size=4)
    with pytest.raises(ValueError):
        R.RecConv2d(4, mode="bicubic")


# ---- the CUDA source's stage schedules, executed on the CPU (tests/emu) ----

def _emu_check(g, dtype, opts, tol):
    from tests.emu import emu

    z, p = g["z"], g["params"]
    y, _ = emu.run(z["x"], p, g["mode"], dtype=dtype, opts=opts)
    assert rel_err(y, z["y"]) < tol
    gr, _ = emu.run(z["x"], p, g["mode"], gy=z["gy"], dtype=dtype, opts=opts)
    assert rel_err(gr["gx"], z["gx"]) < tol
    if g["L"] > 0:
        assert rel_err(gr["down_w"], z["g:down.weight"]) < tol
    for j in range(g["L"] + 1):
        assert rel_err(gr["convs_w"][j], z[f"g:convs.{j}.weight"]) < tol
    if g["bias"]:
        if g["L"] > 0:
            assert rel_err(gr["down_b"], z["g:down.bias"]) < tol
        for j in range(g["L"] + 1):
            assert rel_err(gr["convs_b"][j], z[f"g:convs.{j}.bias"]) < tol


_EMU_CASES = [p for p in recconv_golden_files() if "100x167" not in p]


@pytest.mark.parametrize("path", _EMU_CASES, ids=lambda p: os.path.basename(p)[8:-4])
def test_kernel_schedule_emulated_fp32(path):
    _emu_check(load_recconv_golden(path), 0, (0, 0, 0, 0), TOL_FP32)


@pytest.mark.parametrize("opts", [(1, 32, 0, 1), (8, 4, 2, 0), (3, 32, 1, 1), (2, 64, 0, 0)], ids=str)
def test_kernel_schedule_emulated_forced_tilings(opts):
    ran = 0
    for name in ("m_stage1_28_L3", "m_stage3_7_L1_bias", "k3_33x17_L3_nearest_bias", "tiny_5x3_L4_bias", "m_stage2_14_L2_nearest"):
        try:
            _emu_check(load_recconv_golden(os.path.join(GOLDEN, f"recconv_{name}.npz")), 0, opts, TOL_FP32)
            ran += 1
        except RuntimeError as ex:  # a forced tiling may not fit in shared memory for the bigger planes
            assert "rc=-1" in str(ex)
    assert ran >= 3


def test_kernel_schedule_emulated_bf16_and_fp16():
    for name in ("m_stage2_14_L2_nearest", "m_stage3_7_L1_bias", "k7_40x31_L2"):
        g = load_recconv_golden(os.path.join(GOLDEN, f"recconv_{name}.npz"))
        _emu_check(g, 1, (0, 0, 0, 0), TOL_BF16)
        _emu_check(g, 2, (0, 0, 0, 0), 2e-3)


def test_kernel_schedule_emulated_big_plane_forward_only():
    """100x167 fp32 fits only with the shared raw buffer; its backward is reported as unsupported."""
    from tests.emu import emu

    g = load_recconv_golden(os.path.join(GOLDEN, "recconv_det_odd_100x167_L3.npz"))
    y, plan = emu.run(g["z"]["x"], g["params"], g["mode"])
    assert rel_err(y, g["z"]["y"]) < TOL_FP32
    with pytest.raises(RuntimeError):
        emu.run(g["z"]["x"], g["params"], g["mode"], gy=g["z"]["gy"])


# ---- the TEAM-RESIDENT schedules (wbody.cuh: the fast path), executed on the CPU ----
def _team_check(g, dtype, opts, tol):
    from tests.emu import emu

    z, p = g["z"], g["params"]
    L = g["L"]
    y, plan = emu.run(z["x"], p, g["mode"], dtype=dtype, opts=opts, team=True)
    assert rel_err(y, z["y"]) < tol, ("y", plan)
    gr, plan = emu.run(z["x"], p, g["mode"], gy=z["gy"], dtype=dtype, opts=opts, team=True)
    assert rel_err(gr["gx"], z["gx"]) < tol, ("gx", plan)
    if L > 0:
        assert rel_err(gr["down_w"], z["g:down.weight"]) < tol, ("down_w", plan)
    for j in range(L + 1):
        assert rel_err(gr["convs_w"][j], z[f"g:convs.{j}.weight"]) < tol, (f"convs_w{j}", plan)
    if g["bias"]:
        if L > 0:
            assert rel_err(gr["down_b"], z["g:down.bias"]) < tol
        for j in range(L + 1):
            assert rel_err(gr["convs_b"][j], z[f"g:convs.{j}.bias"]) < tol
    return plan


@pytest.mark.parametrize("path", [p for p in recconv_golden_files() if "100x167" not in p], ids=os.path.basename)
def test_team_schedule_emulated_fp32(path):
    """Forward + backward of the team-resident schedule against the reference fixtures (planner defaults)."""
    _team_check(load_recconv_golden(path), 0, (0, 0, 0, 0, 0), TOL_FP32)


@pytest.mark.parametrize("opts", [(0, 2, 0, 1, 3), (0, 4, 1, 0, 2), (2, 1, 2, 0, 5), (0, 1, 1, 1, 1)],
                         ids=["2warps-noTMA-3sms", "4warps-1team-2sms", "2planes-2teams-5sms", "1team-1sm"])
def test_team_schedule_emulated_forced_tilings(opts):
    """Multi-warp teams, plane batching, few SMs (teams walk over channel groups), cooperative loads."""
    ran = 0
    for name in ("m_stage0_56_L4", "m_stage2_14_L2_nearest", "k3_33x17_L3_nearest_bias", "k7_40x31_L2", "level0_9x11",
                 "nearest_odd_23x29_L2"):
        g = load_recconv_golden(os.path.join(GOLDEN, f"recconv_{name}.npz"))
        if opts[0] and g["C"] % opts[0]:
            continue
        try:
            _team_check(g, 0, opts, TOL_FP32)
            ran += 1
        except RuntimeError as ex:  # the forced tiling does not fit on chip for this shape: the planner says so
            assert "rc=-1" in str(ex)
    assert ran >= 2


def test_team_schedule_emulated_bf16_and_fp16():
    g = load_recconv_golden(os.path.join(GOLDEN, "recconv_m_stage1_28_L3.npz"))
    _team_check(g, 1, (0, 0, 0, 0, 0), TOL_BF16)
    _team_check(g, 2, (0, 0, 0, 0, 0), 2e-3)


def test_team_plan_covers_every_plane_exactly_once():
    """Work split of wplan.h: every (image, channel group) pair is visited by exactly one team, for the BASELINE shapes
    and for few-team grids (emulated through the planner's own iterator in tests/emu)."""
    from tests.emu import emu

    rng = np.random.default_rng(3)
    for (B, C, H, W, L), opts in [((5, 8, 7, 7, 1), (0, 0, 0, 0, 1)), ((3, 6, 9, 11, 2), (2, 1, 2, 0, 1)), ((7, 4, 14, 14, 2), (0, 0, 0, 0, 3))]:
        p = O.RecConvParams.random(C, 5, L, False, rng)
        x = rng.standard_normal((B, C, H, W), dtype=np.float32)
        y, plan = emu.run(x, p, "bilinear", opts=opts, team=True)
        assert rel_err(y, O.forward(x, p, "bilinear")) < TOL_FP32, plan  # a skipped plane would stay zero


def test_tensor_core_plan_selection(native):
    """16-bit activations with k = 5 take the tensor-core forward (compile-time geometry for the RecNeXt stage shapes);
    fp32, k != 5 and tiny planes stay on the FMA kernels; the 16-bit backward of planes >= 20 x 20 is tensor-core too."""
    import torch

    from recnext_b200 import recconv

    d = lambda shape, L, dt, k=5, bwd=False: recconv.plan_describe(shape, k, L, "bilinear", dt, False, bwd)  # noqa: E731
    for shape, L in [((256, 64, 56, 56), 4), ((256, 128, 28, 28), 3), ((256, 256, 14, 14), 2), ((128, 80, 56, 56), 4)]:
        s = d(shape, L, torch.bfloat16)
        assert "tensor-core" in s and "geometry=compile-time" in s, s
    assert "geometry=run-time" in d((2, 128, 100, 168), 3, torch.bfloat16)
    assert "geometry=run-time" in d((4, 64, 56, 56), 4, torch.float16)
    assert "tensor-core" not in d((256, 512, 7, 7), 1, torch.bfloat16)      # smaller than one MMA tile
    assert "tensor-core" not in d((256, 64, 56, 56), 4, torch.float32)
    assert "tensor-core" not in d((4, 64, 56, 56), 4, torch.bfloat16, k=7)
    # backward: tensor-core kernel for 16-bit planes of at least 20 x 20 (smaller ones pack several planes per warp in the FMA kernel)
    assert "bwd tensor-core" in d((4, 64, 56, 56), 4, torch.bfloat16, bwd=True)
    assert "bwd tensor-core" in d((2, 128, 100, 168), 3, torch.bfloat16, bwd=True)       # detection stage 1 fits on chip in 16 bits
    assert "tensor-core" not in d((256, 256, 14, 14), 2, torch.bfloat16, bwd=True)
    assert "tensor-core" not in d((4, 64, 56, 56), 4, torch.float32, bwd=True)
    assert "streamed" in d((2, 64, 200, 336), 4, torch.bfloat16, bwd=True)               # detection stage 0: level by level
    # teams fill the SM: at least 8 warps for every stage shape
    import re
    for shape, L in [((256, 64, 56, 56), 4), ((256, 128, 28, 28), 3), ((256, 256, 14, 14), 2), ((2, 128, 100, 168), 3)]:
        s = d(shape, L, torch.bfloat16)
        assert int(re.search(r"threads=(\d+)", s).group(1)) >= 256, s
        assert int(re.search(r"smem=(\d+)", s).group(1)) <= 227 * 1024, s
