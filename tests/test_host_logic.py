"""CPU tests of the product's host logic: the C-ABI library loads, exports every symbol the header
declares, its host-side pattern generators agree with the oracle, and compute entry points fail
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import slr_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = slr_b200.capi()
    header = open(os.path.join(ROOT, "include", "slr_b200.h")).read()
    declared = set(re.findall(r"\b(slr_[a-z0-9_]+)\s*\(", header))
    declared -= {"slr_status", "slr_mode"}
    assert declared, "no declarations parsed"
    assert declared == set(slr_b200.EXPORTS), declared ^ set(slr_b200.EXPORTS)
    for name in sorted(declared):
        assert getattr(lib, name) is not None, name


def test_version_and_error_strings():
    lib = slr_b200.capi()
    assert b"sm_100a" in lib.slr_version()
    assert isinstance(lib.slr_last_error(), bytes)


@pytest.mark.parametrize("W,H,epi", [(1280, 8, True), (640, 6, True), (80, 72, False), (1280, 1024, False)])
def test_gray_patterns_match_oracle(oracle, W, H, epi):
    a = slr_b200.generate_gray_patterns(W, H, epi)
    b = oracle.generate_gray(W, H, epi)
    assert a.shape == b.shape and (a == b).all()


@pytest.mark.parametrize("W,H", [(1280, 4), (1024, 3), (912, 2)])
def test_mf_patterns_match_oracle(oracle, W, H):
    a = slr_b200.generate_mf_patterns(W, H)
    b = oracle.generate_mf(W, H)
    assert (a == b).all()


def test_gray_num_bits_matches_oracle(oracle):
    for n in (2, 3, 480, 640, 800, 1024, 1280, 2048, 3000, 4096):
        assert slr_b200.gray_num_bits(n) == oracle.gray_num_bits(n)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = slr_b200.capi()
    h = C.c_void_p()
    st = lib.slr_create(C.byref(h), 0, 64, 16, 1)
    assert st != 0 and not h.value
    assert b"no CPU fallback" in lib.slr_last_error()
    with pytest.raises(slr_b200.SlrError):
        slr_b200.Engine(64, 16)


def test_bad_arguments_are_rejected():
    lib = slr_b200.capi()
    assert lib.slr_create(None, 0, 64, 16, 1) != 0
    h = C.c_void_p()
    assert lib.slr_create(C.byref(h), 0, 0, 16, 1) != 0
    assert lib.slr_generate_mf_patterns(None, 16, 16) != 0
    assert lib.slr_mf_decode(None, None, 1, 3, 4, 40, 0, None, None) != 0


def test_numpy_synth_shapes_and_determinism():
    from slr_b200 import synth
    a = synth.synth_mf(64, 16, seed=3)
    b = synth.synth_mf(64, 16, seed=3)
    assert a.shape == (2, 14, 16, 64) and (a == b).all()
    g = synth.synth_gray(64, 16, seed=3)
    assert g.shape == (2, 2 + 2 * 6, 16, 64)
    g2 = synth.synth_gray(64, 16, seed=3, rows=True)
    assert g2.shape == (2, 2 + 2 * 6 + 2 * 4, 16, 64)


def test_oracle_pipeline_on_numpy_synth(oracle):
    """Host-only end-to-end sanity of the test inputs: integer-disparity scenes give a dense cloud."""
    from slr_b200 import synth
    W, H = 256, 8
    st = synth.synth_mf(W, H, seed=5, integer_disparity=True)
    cams, Q = slr_b200.synthetic_rig(W, H)
    xyz, valid, mk, n = oracle.run_mf(st, cams, Q)
    assert n == valid.sum() and n > 0.5 * W * H
    assert np.isfinite(xyz[valid == 1]).all() and np.isnan(xyz[valid == 0]).all()
