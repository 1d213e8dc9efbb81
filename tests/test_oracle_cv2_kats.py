"""Oracle's OpenCV primitives vs real cv2 outputs (tests/golden/cv2_kats.npz, oracle/pin_cv2.py).

Bit-exact: these pin the float/double evaluation rules of SURVEY.md §8(c)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_py as O


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "cv2_kats.npz"))


@pytest.fixture(scope="module")
def lib():
    return O.lib("canonical")


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_fast_atan2(kat, lib):
    out = np.array([lib.ocv_fastAtan2(float(y), float(x)) for y, x in zip(kat["atan_y"], kat["atan_x"])], np.float32)
    assert np.array_equal(bits(out), bits(kat["atan_out"]))


def test_abt(kat, lib):
    for key, alpha in (("abt_pos", 1.0), ("abt_neg", -1.0)):
        for a, b, ref in zip(kat["abt_A"], kat["abt_B"], kat[key]):
            d = np.zeros((3, 3), np.float32)
            lib.ocv_mul33_ABt(O.fptr(_f(a)), O.fptr(_f(b)), alpha, O.fptr(d))
            assert np.array_equal(bits(d), bits(ref))


def test_mul33(kat, lib):
    for a, b, ref in zip(kat["mm_A"], kat["mm_B"], kat["mm_out"]):
        d = np.zeros((3, 3), np.float32)
        lib.ocv_mul33(O.fptr(_f(a)), O.fptr(_f(b)), O.fptr(d))
        assert np.array_equal(bits(d), bits(ref))


def test_mul33_vec(kat, lib):
    for a, x, c, al, ref, ref0 in zip(kat["mv_A"], kat["mv_x"], kat["mv_c"], kat["mv_alpha"], kat["mv_out"],
                                      kat["mv_out_noc"]):
        d = np.zeros(3, np.float32)
        lib.ocv_mul33_vec(O.fptr(_f(a)), O.fptr(_f(x)), float(al), O.fptr(_f(c)), 1.0, O.fptr(d))
        assert np.array_equal(bits(d), bits(ref.reshape(3)))
        lib.ocv_mul33_vec(O.fptr(_f(a)), O.fptr(_f(x)), float(al), None, 0.0, O.fptr(d))
        assert np.array_equal(bits(d), bits(ref0.reshape(3)))


def test_dot3(kat, lib):
    out = np.array([lib.ocv_dot3_d(O.fptr(_f(a)), O.fptr(_f(x)), float(al))
                    for a, x, al in zip(kat["dot_a"], kat["mv_x"], kat["mv_alpha"])], np.float32)
    assert np.array_equal(bits(out), bits(kat["dot_out"]))


def test_dotn(kat, lib):
    for L, J, r, neg, jtj in zip(kat["jn_len"], kat["jn_J"], kat["jn_r"], kat["jn_neg"], kat["jn_jtj"]):
        a = lib.ocv_dotn_d(O.fptr(_f(J)), O.fptr(_f(r)), int(L), -1.0)
        b = lib.ocv_dotn_d(O.fptr(_f(J)), O.fptr(_f(J)), int(L), 1.0)
        assert np.float32(a).view(np.uint32) == np.float32(neg).view(np.uint32)
        assert np.float32(b).view(np.uint32) == np.float32(jtj).view(np.uint32)


def test_inv33(kat, lib):
    for a, ref in zip(kat["inv_A"], kat["inv_out"]):
        d = np.zeros((3, 3), np.float32)
        lib.ocv_inv33(O.fptr(_f(a)), O.fptr(d))
        assert np.array_equal(bits(d), bits(ref))


def test_solve33(kat, lib):
    for a, b, ref in zip(kat["solve_A"], kat["solve_B"], kat["solve_out"]):
        d = np.zeros((3, 3), np.float32)
        lib.ocv_solve33_lu(O.fptr(_f(a)), O.fptr(_f(b)), O.fptr(d))
        assert np.array_equal(bits(d), bits(ref))


def test_mul44_vec(kat, lib):
    for a, x, ref in zip(kat["m4_A"], kat["m4_x"], kat["m4_out"]):
        d = np.zeros(4, np.float32)
        lib.ocv_mul44_vec(O.fptr(_f(a)), O.fptr(_f(x)), O.fptr(d))
        assert np.array_equal(bits(d), bits(ref.reshape(4)))


def test_scharr_numpy_matches_cv2(kat):
    """The numpy plane producer used when cv2 is absent reproduces cv2.Scharr/32 exactly."""
    from sdmb200 import synth
    im = kat["sch_img"]
    p = np.pad(im.astype(np.int32), 1, mode="reflect")
    gx = (3 * (p[:-2, 2:] - p[:-2, :-2]) + 10 * (p[1:-1, 2:] - p[1:-1, :-2]) + 3 * (p[2:, 2:] - p[2:, :-2])) / 32.0
    gy = (3 * (p[2:, :-2] - p[:-2, :-2]) + 10 * (p[2:, 1:-1] - p[:-2, 1:-1]) + 3 * (p[2:, 2:] - p[:-2, 2:])) / 32.0
    assert np.array_equal(gx.astype(np.float32), kat["sch_gx"])
    assert np.array_equal(gy.astype(np.float32), kat["sch_gy"])
    g, t = synth.gradient_planes(im)
    assert g.shape == im.shape and t.shape == im.shape


def test_numpy_fast_atan2_matches_cv2(kat):
    """the generator's scalar-form phase (synth._fast_atan2_deg, also what the device plane producer computes)
    == cv2.fastAtan2 on the known-answer inputs"""
    from sdmb200 import synth
    out = synth._fast_atan2_deg(kat["atan_y"].astype(np.float32), kat["atan_x"].astype(np.float32))
    assert np.array_equal(bits(out), bits(kat["atan_out"]))
