"""Drop-in proof: the product's ModCodec adapter (odr-dabmod_b200/adapter/
B200OfdmChain), compiled against the UNMODIFIED reference headers, runs inside
the reference's own Flowgraph (oracle/adapter_harness.cpp) and is compared with
the all-reference graph (oracle/ref_harness.cpp) on the same input.

Needs the prebuilt oracle/_ref/*.so (they travel to the GPU box with the snapshot).
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import rel_rms, write_poly_file
from oracle import refwrap

ADP_LIB = os.path.join(os.path.dirname(refwrap.REF_LIB), "libdabmod_adapter.so")
have = os.path.exists(ADP_LIB) and refwrap.available()
TOL = 2e-6


def adp():
    L = ctypes.CDLL(ADP_LIB)
    L.adp_create.restype = ctypes.c_void_p
    L.adp_create.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.adp_process.restype = ctypes.c_long
    L.adp_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    L.adp_set_parameter.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    L.adp_get_parameter.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    L.adp_destroy.argtypes = [ctypes.c_void_p]
    L.adp_last_error.restype = ctypes.c_char_p
    L.adp_create2.restype = ctypes.c_void_p
    L.adp_create2.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    L.adp_last_metadata_fct.argtypes = [ctypes.c_void_p]
    L.adp_tii_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    L.adp_setting.restype = ctypes.c_double
    L.adp_setting.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    return L


@pytest.mark.skipif(not have, reason="reference libraries not built")
def test_adapter_library_links():
    """CPU check: the adapter library loads (so it resolved every dabmod_b200_* symbol it binds)."""
    L = adp()
    for sym in ("adp_create", "adp_process", "adp_set_parameter", "adp_destroy"):
        assert hasattr(L, sym)


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="reference libraries not built")
@pytest.mark.parametrize("case", ["c1", "c2", "c3", "tm2_s16_tii", "tm1_fixed", "tm4_fixed_window"])
def test_adapter_in_reference_flowgraph(rng, tmp_path, case):
    L = adp()
    kw = {"c1": dict(mode=1), "c2": dict(mode=1, fir_taps_file="default"),
          "c3": dict(mode=1, fir_taps_file="default", output_rate=8192000, normalise=1.0 / 46000.0),
          "tm2_s16_tii": dict(mode=2, tii=(3, 20, 0), digital_gain=0.8, fmt="s16"),
          "tm1_fixed": dict(mode=1, fixed_point=True, tii=(5, 8, 0)),
          "tm4_fixed_window": dict(mode=4, fixed_point=True, window_overlap=20)}[case]
    fixed = bool(kw.get("fixed_point"))
    if case == "c3":
        p = str(tmp_path / "poly.coef")
        write_poly_file(p, [1.0, 0.05, -0.02, 0.0, 0.0], [0.0, 0.1, -0.05, 0.0, 0.0])
        kw["poly_coef_file"] = p
        kw["poly_threads"] = 1
    mode = kw["mode"]
    dt = np.int16 if kw.get("fmt") == "s16" or fixed else np.complex64
    bits = rng.integers(0, 256, (3, refwrap.TF_BYTES[mode]), dtype=np.uint8)
    ref = refwrap.RefChain(**kw)
    want = ref.run(bits, dtype=dt)
    h = L.adp_create(ctypes.byref(ref._cfg), 0)
    assert h, L.adp_last_error().decode()
    out = np.empty(64 << 20, np.uint8)
    for i in range(3):
        n = L.adp_process(h, bits[i].ctypes.data, bits[i].size, out.ctypes.data, out.size)
        assert n > 0, L.adp_last_error().decode()      # no priming latency: every call returns its TF
        got = out[:n].view(dt)
        assert got.size == want[i].size
        if fixed:
            assert np.array_equal(got, want[i])          # the fixed-point engine is bit-exact
        elif dt is np.int16:
            d = np.abs(got.astype(np.int32) - want[i].astype(np.int32))
            assert d.max() <= 1 and np.count_nonzero(d) < 0.01 * d.size
        else:
            assert rel_rms(got, want[i]) < TOL
    # wrong input size throws inside process() like the reference blocks do -> harness reports -1
    assert L.adp_process(h, bits[0].ctypes.data, 100, out.ctypes.data, out.size) == -1
    if fixed:
        assert L.adp_set_parameter(h, b"digital", b"0.5") == -1      # no GainControl in the fixed-point chain
        L.adp_destroy(h)
        return
    # remote control through the RemoteControllable interface
    assert L.adp_set_parameter(h, b"digital", b"0.5") == 0
    buf = ctypes.create_string_buffer(64)
    assert L.adp_get_parameter(h, b"digital", buf, 64) == 0 and float(buf.value) == 0.5
    assert L.adp_set_parameter(h, b"mode", b"bogus") == -1
    L.adp_destroy(h)


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="reference libraries not built")
@pytest.mark.parametrize("depth", [1, 4])
def test_adapter_pipeline_delays_data_and_metadata_together(rng, depth):
    """The N-TF generalisation of PipelinedModCodec (src/ModPlugin.cpp:90-128): call i returns TF i - depth, the
    first `depth` calls end the flowgraph iteration (0), every depth-th call runs one batch of `depth` TFs, and the
    frame timestamps reach OutputMemory with the TF they belong to."""
    L = adp()
    kw = dict(mode=1, fir_taps_file="default", tii=(3, 20, 0))      # TII: the batch must keep the every-second-TF toggle
    n = 3 * depth + 2
    bits = rng.integers(0, 256, (n, refwrap.TF_BYTES[1]), dtype=np.uint8)
    ref = refwrap.RefChain(**kw)
    want = ref.run(bits)
    h = L.adp_create2(ctypes.byref(ref._cfg), 0, depth)
    assert h, L.adp_last_error().decode()
    out = np.empty(16 << 20, np.uint8)
    for i in range(n):
        nb = L.adp_process(h, bits[i].ctypes.data, bits[i].size, out.ctypes.data, out.size)
        if i < depth:
            assert nb == 0                                           # priming: nothing downstream runs
            continue
        assert nb > 0, L.adp_last_error().decode()
        assert rel_rms(out[:nb].view(np.complex64), want[i - depth]) < TOL, i
        assert L.adp_last_metadata_fct(h) == i - depth               # the timestamp of the TF that came out
    L.adp_destroy(h)


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="reference libraries not built")
def test_adapter_remote_control_writes_back_and_tii_controllable(rng):
    """Remote-control changes land in the mod_settings_t the chain was built from (the reference blocks hold
    references into it, e.g. src/GainControl.h:72-77, so a rebuilt modulator keeps them), and TII is its own
    controllable named "tii" with the reference's parameter names (src/TII.cpp:106-127)."""
    L = adp()
    ref = refwrap.RefChain(mode=1)
    h = L.adp_create(ctypes.byref(ref._cfg), 0)
    assert h, L.adp_last_error().decode()
    for name, value, want in [("digital", "0.5", 0.5), ("var", "3", 3.0), ("mode", "max", 1.0), ("windowlen", "12", 12.0),
                              ("cfr", "1", 1.0), ("clip", "70", 70.0), ("errorclip", "0.2", 0.2)]:
        assert L.adp_set_parameter(h, name.encode(), value.encode()) == 0, L.adp_last_error().decode()
        assert abs(L.adp_setting(h, name.encode()) - want) < 1e-6, name
    for name, value in [("comb", "5"), ("pattern", "33"), ("enable", "1")]:
        assert L.adp_tii_set(h, name.encode(), value.encode()) == 0, L.adp_last_error().decode()
        assert L.adp_setting(h, ("tii." + name).encode()) == float(value)
    assert L.adp_tii_set(h, b"comb", b"99") == -1                    # TII comb not valid! (src/TII.cpp:346-350)
    assert L.adp_setting(h, b"tii.comb") == 5.0                      # a refused value does not reach the settings
    buf = ctypes.create_string_buffer(256)
    assert L.adp_get_parameter(h, b"tii.pattern", buf, 256) == 0 and buf.value == b"33"
    L.adp_destroy(h)
