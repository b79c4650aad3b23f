"""Boundary shape B1 at block level: oracle/_ref/libdabmod_b1.so is the reference's own hot-path harness
(oracle/ref_harness.cpp: the blocks constructed and wired like DabModulator does, src/DabModulator.cpp:144-417) linked with
the product's adapter/B200Blocks.cpp IN PLACE OF the sixteen reference translation units.  The very same harness linked
with the reference's own files is libdabmod_ref.so -- the checker of every other parity test -- so each case below
feeds both with the same transmission frames and compares what falls out of OutputMemory, call by call (the pipelined
blocks delay both alike)."""
import os

import numpy as np
import pytest

from conftest import rel_rms, write_poly_file
from oracle import refwrap

have = refwrap.available() and refwrap.available("b1")
needs_libs = pytest.mark.skipif(not have, reason="oracle/_ref libraries not built (make -C oracle ref b1)")
pytestmark = [needs_libs]


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="needs a host without a GPU")
def test_substituted_blocks_have_no_cpu_fallback(rng):
    """CPU: the library links (every reference symbol the harness needs is defined by B200Blocks.cpp), the graph
    builds, the tokens pass the pipelined blocks (empty calls while they prime) -- and the first frame that reaches
    OutputMemory fails with the C ABI's error instead of being computed on the host."""
    b1 = refwrap.RefChain(variant="b1", mode=2, fir_taps_file="default")
    assert b1.latency == 2                                     # GainControl + FIRFilter, like the reference graph
    bits = rng.integers(0, 256, refwrap.TF_BYTES[2], dtype=np.uint8)
    for _ in range(b1.latency):
        assert b1.feed(bits).size == 0
    with pytest.raises(RuntimeError, match="no CUDA device"):
        b1.feed(bits)
    b1.close()

CASES = {
    "tm1_var": dict(mode=1),
    "tm2_max": dict(mode=2, gain_mode="max", digital_gain=0.7),
    "tm3_fix": dict(mode=3, gain_mode="fix"),
    "tm4_var3": dict(mode=4, gain_variance=3.0, normalise=0.5),
    "tm1_fir": dict(mode=1, fir_taps_file="default"),
    "tm1_tii": dict(mode=1, tii=(7, 33, 0)),
    "tm2_tii_old": dict(mode=2, tii=(2, 11, 1)),
    "tm1_window": dict(mode=1, window_overlap=64),
    "tm1_cfr": dict(mode=1, cfr=(90.0, 0.05), gain_mode="var"),
    "tm1_ciceq": dict(mode=1, clock_rate=128000000, output_rate=2048000),
    "tm1_res8": dict(mode=1, fir_taps_file="default", output_rate=8192000),
    "tm1_res10_poly": dict(mode=1, fir_taps_file="default", output_rate=10000000, normalise=1.0 / 46000.0, poly="poly"),
    "tm2_lut": dict(mode=2, normalise=1.0 / 46000.0, poly="lut"),
    "tm1_s16": dict(mode=1, normalise=32767.0 / 46000.0, fmt="s16"),
    "tm4_u8": dict(mode=4, normalise=127.0 / 46000.0, fmt="u8"),
    "tm2_s8_down": dict(mode=2, normalise=127.0 / 46000.0, fmt="s8", output_rate=1536000),
    "tm1_fixed_tii": dict(mode=1, fixed_point=True, tii=(5, 8, 0)),
    "tm3_fixed_window": dict(mode=3, fixed_point=True, window_overlap=12),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_substituted_blocks_against_the_reference_blocks(rng, tmp_path, case):
    kw = dict(CASES[case])
    poly = kw.pop("poly", None)
    if poly == "poly":
        kw["poly_coef_file"] = str(tmp_path / "poly.coef")
        write_poly_file(kw["poly_coef_file"], [1.0, 0.05, -0.02, 0.0, 0.0], [0.0, 0.1, -0.05, 0.0, 0.0])
    elif poly == "lut":
        kw["poly_coef_file"] = str(tmp_path / "lut.coef")
        with open(kw["poly_coef_file"], "w") as f:
            f.write("2\n%r\n" % (2 ** 32 / 1.5) + "\n".join("%r" % (1.0 - 0.004 * i) for i in range(32)) + "\n")
    fixed = bool(kw.get("fixed_point"))
    fmt = kw.get("fmt")
    dt = np.int16 if fixed or fmt == "s16" else np.uint8 if fmt == "u8" else np.int8 if fmt == "s8" else np.complex64
    n_tf = 5
    bits = rng.integers(0, 256, (n_tf, refwrap.TF_BYTES[kw["mode"]]), dtype=np.uint8)
    ref = refwrap.RefChain(**kw)
    b1 = refwrap.RefChain(variant="b1", **kw)
    assert ref.latency == b1.latency
    for i in range(n_tf):
        want, got = ref.feed(bits[i]), b1.feed(bits[i])
        assert got.size == want.size, (case, i, got.size, want.size)        # incl. the empty calls while the pipelines prime
        if want.size == 0:
            continue
        a, b = got.view(dt), want.view(dt)
        if fixed:
            assert np.array_equal(a, b), (case, i)
        elif poly == "lut":
            # a sample whose magnitude lands within float rounding of a bin edge may pick the neighbouring table entry
            assert np.count_nonzero(np.abs(a - b) > 1e-5 * np.abs(b).max()) <= 8, (case, i)
        elif dt is np.complex64:
            assert rel_rms(a, b) < 2e-6, (case, i, rel_rms(a, b))
        else:
            d = np.abs(a.astype(np.int32) - b.astype(np.int32))
            assert d.max() <= 1 and np.count_nonzero(d) < 0.01 * d.size, (case, i, int(d.max()), np.count_nonzero(d))
    if case == "tm1_cfr":
        # the remote-control read-outs of the substituted OfdmGenerator ("ofdm" controllable)
        assert b1.get_param("cfr") == ref.get_param("cfr")
        assert b1.get_param("clip") == ref.get_param("clip")
        # the GPU runs a frame when its token reaches OutputMemory, i.e. `latency` calls after the reference's
        # OfdmGenerator saw it: the statistics cover the same frames once those calls have been made
        for _ in range(b1.latency):
            b1.feed(bits[-1])
        assert b1.get_param("clip_stats") == ref.get_param("clip_stats")
        assert b1.get_param("papr") == ref.get_param("papr")
    ref.close()
    b1.close()


@pytest.mark.gpu
def test_two_graphs_side_by_side(rng):
    """Two graphs built before either runs (different modes), frames interleaved: the token keeps them apart."""
    a = refwrap.RefChain(variant="b1", mode=1, fir_taps_file="default")
    b = refwrap.RefChain(variant="b1", mode=2, tii=(2, 11, 0))
    ra = refwrap.RefChain(mode=1, fir_taps_file="default")
    rb = refwrap.RefChain(mode=2, tii=(2, 11, 0))
    for i in range(4):
        xa = rng.integers(0, 256, refwrap.TF_BYTES[1], dtype=np.uint8)
        xb = rng.integers(0, 256, refwrap.TF_BYTES[2], dtype=np.uint8)
        for got, want in ((a.feed(xa), ra.feed(xa)), (b.feed(xb), rb.feed(xb))):
            assert got.size == want.size
            if want.size:
                assert rel_rms(got.view(np.complex64), want.view(np.complex64)) < 2e-6, i
    for c in (a, b, ra, rb):
        c.close()
