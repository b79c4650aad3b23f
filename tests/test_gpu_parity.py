"""Parity of the CUDA path (through the C ABI) against the oracle.

Tolerance: BASELINE.json north_star -- output I/Q within 1e-5 relative RMS of the
reference CPU path.  We assert 2e-6 (5x tighter) for float32 outputs; integer
output formats may differ by one count where a float lands on a truncation edge.
"""
import numpy as np
import pytest

import dabmod_loader
from conftest import rel_rms
from oracle import oracle

pytestmark = pytest.mark.gpu

TOL = 2e-6


@pytest.fixture(scope="module")
def dm():
    return dabmod_loader.load()


def bits_for(rng, mode, n):
    return rng.integers(0, 256, (n, oracle.mode_params(mode).tf_bytes), dtype=np.uint8)


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_native_rate_all_modes(dm, rng, mode):
    """BASELINE configs 1 and 4: no FIR/resampler, gain var, complexf."""
    bits = bits_for(rng, mode, 3)
    ora = oracle.OracleChain(mode=mode).run(bits)
    mod = dm.Modulator(mode=mode, max_batch=3)
    out = mod.process_batch(bits)
    assert out.shape == (3, oracle.mode_params(mode).tf_samples)
    for i in range(3):
        assert rel_rms(out[i], ora[i]) < TOL
    # single-TF entry point gives the same bytes as the batch entry point
    mod.reset()
    one = mod.process(bits[0])
    assert np.array_equal(one.view(np.uint32), out[0].view(np.uint32))
    assert mod.last_launch_count == 1


@pytest.mark.parametrize("gain_mode,dg,norm,var", [("fix", 1.0, 1.0, 4.0), ("max", 0.8, 1.0, 4.0),
                                                    ("var", 0.7, 1.0 / 46000.0, 3.0)])
@pytest.mark.parametrize("mode", [1, 3])
def test_gain_modes(dm, rng, mode, gain_mode, dg, norm, var):
    bits = bits_for(rng, mode, 2)
    kw = dict(mode=mode, gain_mode=gain_mode, digital_gain=dg, normalise=norm, gain_variance=var)
    ora = oracle.OracleChain(**kw).run(bits)
    out = dm.Modulator(max_batch=2, **kw).process_batch(bits)
    for i in range(2):
        assert rel_rms(out[i], ora[i]) < TOL


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_fir_default_taps(dm, rng, mode):
    """BASELINE config 2: FIRFilter with the built-in taps."""
    bits = bits_for(rng, mode, 2)
    taps = oracle.fir_default_taps()
    assert np.array_equal(taps, dm.default_fir_taps())
    ora = oracle.OracleChain(mode=mode, fir_taps=taps).run(bits)
    mod = dm.Modulator(mode=mode, fir_taps="default", max_batch=2)
    out = mod.process_batch(bits)
    for i in range(2):
        assert rel_rms(out[i], ora[i]) < TOL
    assert mod.last_launch_count == 2      # symbol kernel + FIR kernel (two frames: too few for the fused TM I kernel)


@pytest.mark.parametrize("kw", [dict(), dict(fmt="s16", digital_gain=0.7), dict(gain_mode="max"),
                                dict(output_rate=4096000), dict(poly=[1.0, 0.05, -0.02, 0.003, 0.0, 0.0, 0.1, -0.05, 0.01, 0.0],
                                                                normalise=1.0 / 46000.0)],
                         ids=["complexf", "s16", "gain_max", "resampler_after", "poly"])
def test_fir_symbol_kernel(dm, rng, kw):
    """k_fir_sym (TM I, default taps, symbol kernel in its compact layout: the cyclic-prefix outputs are copies of the
    tail outputs, the prefix itself is never stored) produces the same BITS as the sample-stream k_fir, and both
    agree with the oracle."""
    bits = bits_for(rng, 1, 3)
    taps = oracle.fir_default_taps()
    a = dm.Modulator(mode=1, fir_taps=taps, max_batch=3, **kw)
    a.set_param("profile", 1)
    a.set_param("fir_kernel", 2)                 # (3, the default, puts the filter into the symbol kernel: next test)
    ya = a.process_batch(bits)
    # complexf straight out of the FIR: the persistent TMA-fed variant; with an epilogue (format, predistortion): k_fir_sym
    raw = "fmt" not in kw and "poly" not in kw or "output_rate" in kw
    assert ("k_fir_tma" if raw else "k_fir_sym") in [k for k, _ in a.kernel_times()]
    if raw:
        c = dm.Modulator(mode=1, fir_taps=taps, max_batch=3, **kw)
        c.set_param("fir_kernel", 1)
        c.set_param("profile", 1)
        yc = c.process_batch(bits)
        assert "k_fir_sym" in [k for k, _ in c.kernel_times()]
        assert np.array_equal(ya.view(np.uint8), yc.view(np.uint8))
    b = dm.Modulator(mode=1, fir_taps=taps, max_batch=3, **kw)
    b.set_param("fir_kernel", 0)
    b.set_param("profile", 1)
    yb = b.process_batch(bits)
    assert "k_fir" in [k for k, _ in b.kernel_times()]
    assert ya.dtype == yb.dtype and np.array_equal(ya.view(np.uint8), yb.view(np.uint8))
    if "fmt" not in kw:
        want = oracle.OracleChain(mode=1, fir_taps=taps, **kw).run(bits)
        for i in range(3):
            assert rel_rms(ya[i], want[i]) < TOL, i
    # one TF at a time gives the same bits again (the last symbol's window ends in zeros, not in the next TF)
    a.reset()
    for i in range(3):
        assert np.array_equal(a.process(bits[i]).view(np.uint8), ya[i].view(np.uint8)), i


@pytest.mark.parametrize("n_tf,kw", [(1, {}), (3, {}), (7, dict(gain_mode="max")), (40, {}), (64, dict(output_rate=8192000)),
                                     (3, dict(tii=(3, 20))), (40, dict(tii=(3, 20))), (33, dict(tii=(7, 5, 1), gain_mode="max")),
                                     (200, {}), (333, dict(gain_mode="fix")),
                                     (256, dict(output_rate=8192000))])
def test_fir_fused_into_the_symbol_kernel(dm, rng, n_tf, kw):
    """`fir_kernel` = 3: the 45-tap FIR runs inside k_symbols_w on the staged symbol (symbols_warp.cuh, FUSE); the tail
    of a symbol is finished one iteration later from carried samples, a warp whose range ends inside a TF assembles
    one symbol more.  Same operands in the same order as k_fir: the same BITS, whatever the batch size (few TFs: most
    warps idle or with one symbol; many: ranges that end anywhere in a TF)."""
    bits = bits_for(rng, 1, n_tf)
    taps = oracle.fir_default_taps()
    ref = dm.Modulator(mode=1, fir_taps=taps, max_batch=n_tf, **kw)
    ref.set_param("fir_kernel", 0)
    want = ref.process_batch(bits)
    f = dm.Modulator(mode=1, fir_taps=taps, max_batch=n_tf, **kw)
    f.set_param("fir_kernel", 4)                     # (3, the default, only fuses from 8 symbols per warp on)
    f.set_param("profile", 1)
    got = f.process_batch(bits)
    names = [k for k, _ in f.kernel_times()]
    if "tii" in kw and n_tf < 16:
        assert "k_symbols_w_fir" not in names          # a few TII frames: the general kernel + k_fir
    elif "tii" in kw:
        # (k_symbols + k_fir on ONE frame make the stream's filtered TII symbol; the host entry point works in slices,
        # a last slice below 16 frames takes the general kernels)
        assert "k_symbols_w_fir" in names and "k_tii_fill" in names
    else:
        assert "k_symbols_w_fir" in names and not any(k.startswith("k_fir") for k in names)
        # the default (3) decides per launch -- the host entry point works in slices of ~30 TFs -- and gives the same bits
        d = dm.Modulator(mode=1, fir_taps=taps, max_batch=n_tf, **kw)
        assert np.array_equal(d.process_batch(bits).view(np.uint8), want.view(np.uint8))
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    # a second call on the same handle, and frame by frame
    assert np.array_equal(f.process_batch(bits).view(np.uint8), want.view(np.uint8)) or "output_rate" in kw or "tii" in kw
    if "tii" in kw and n_tf >= 16:
        # the TII toggle runs on across calls (TII.cpp:225-242): an odd number of frames flips which frames carry it
        ref.reset()
        f.reset()
        a = np.concatenate([ref.process_batch(bits[:17]), ref.process_batch(bits[17:])])
        b = np.concatenate([f.process_batch(bits[:17]), f.process_batch(bits[17:])])
        # (against each other: with TII the symbol kernel itself depends on the number of frames in a launch -- the
        # warp kernels from 16 frames on -- and the two families agree to 1e-7, not to the bit)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
        assert rel_rms(b.reshape(-1), want.reshape(-1)) < 1e-6
    if "output_rate" not in kw and "tii" not in kw and n_tf <= 7:
        f.reset()
        for i in range(n_tf):
            assert np.array_equal(f.process(bits[i]).view(np.uint8), want[i].view(np.uint8)), i


@pytest.mark.parametrize("ntaps", [1, 2, 16, 17, 33, 64, 97, 128])
def test_fir_tap_counts(dm, rng, ntaps):
    bits = bits_for(rng, 2, 1)
    taps = (rng.standard_normal(ntaps) / np.sqrt(ntaps)).astype(np.float32)
    ora = oracle.OracleChain(mode=2, fir_taps=taps).run(bits)
    out = dm.Modulator(mode=2, fir_taps=taps).process_batch(bits)
    assert rel_rms(out[0], ora[0]) < TOL


@pytest.mark.parametrize("mode,ntaps,fmt", [(2, 129, None), (2, 300, None), (4, 1000, "s16"), (1, 2049, None)])
def test_fir_long_filters(dm, rng, mode, ntaps, fmt):
    """FIRFilter::load_filter_taps takes any tap count (src/FIRFilter.cpp:95-141): above 128 taps the filter runs
    from a device tap table (k_fir_long), also longer than a whole symbol; the window still ends in zeros at the TF end,
    and the RC `taps` parameter switches between the short and the long kernel."""
    bits = bits_for(rng, mode, 2)
    taps = (rng.standard_normal(ntaps) / np.sqrt(ntaps)).astype(np.float32)
    kw = dict(fmt=fmt, digital_gain=0.5) if fmt else {}
    ora = oracle.OracleChain(mode=mode, fir_taps=taps, **kw).run(bits)
    mod = dm.Modulator(mode=mode, fir_taps=taps, max_batch=2, **kw)
    mod.set_param("profile", 1)
    out = mod.process_batch(bits)
    assert "k_fir_long" in [k for k, _ in mod.kernel_times()]
    for i in range(2):
        if fmt:
            d = np.abs(out[i].astype(np.int32) - ora[i].astype(np.int32))
            assert d.max() <= 1 and np.count_nonzero(d) < 0.01 * d.size
        else:
            assert rel_rms(out[i], ora[i]) < TOL, i
    if not fmt:
        short = (rng.standard_normal(20) / 4).astype(np.float32)
        mod.set_param("taps", "20 " + " ".join("%.9g" % t for t in short))
        mod.reset()
        a = mod.process_batch(bits)
        assert "k_fir" in [k for k, _ in mod.kernel_times()] and "k_fir_long" not in [k for k, _ in mod.kernel_times()]
        want = oracle.OracleChain(mode=mode, fir_taps=short).run(bits)
        assert rel_rms(a[0], want[0]) < TOL
        mod.set_param("taps", "%d " % ntaps + " ".join("%.9g" % t for t in taps))
        mod.reset()
        assert np.array_equal(mod.process_batch(bits).view(np.uint32), out.view(np.uint32))


def test_input_edge_patterns(dm):
    """All-zero, all-one and alternating bit blocks (every bit pattern is valid QPSK input)."""
    m = oracle.mode_params(1)
    pats = np.stack([np.zeros(m.tf_bytes, np.uint8), np.full(m.tf_bytes, 0xFF, np.uint8),
                     np.tile(np.array([0xAA, 0x55], np.uint8), m.tf_bytes // 2)])
    ora = oracle.OracleChain(mode=1).run(pats)
    out = dm.Modulator(mode=1, max_batch=3).process_batch(pats)
    for i in range(3):
        assert rel_rms(out[i], ora[i]) < TOL


def test_errors(dm, rng):
    mod = dm.Modulator(mode=1, max_batch=2)
    with pytest.raises(dm.DabModError):
        mod.process(np.zeros(100, np.uint8))          # wrong input size
    with pytest.raises(dm.DabModError):
        mod.process_batch(bits_for(rng, 1, 3))        # exceeds max_batch
    with pytest.raises(dm.DabModError):
        dm.Modulator(mode=7)
    with pytest.raises(dm.DabModError):
        mod.set_param("nonexistent", "1")
    # empty batch is a no-op
    assert mod.process_batch(np.zeros((0, mod.tf_in_bytes), np.uint8)).size == 0


# ---------------------------------------------------------------------------
# Resampler (Resampler.cpp:131-195): FFT overlap-add with state across TFs
# ---------------------------------------------------------------------------
RES_CASES = [(1, 8192000), (1, 10000000), (2, 4096000), (1, 1536000), (4, 2500000), (3, 2400000), (2, 3200000),
             (1, 4000000), (1, 6000000), (1, 8000000),
             # output transform sizes with a prime factor above 7 (generic O(R) pass): 5632 = 2^9 * 11, 1664 = 2^7 * 13,
             # and down-sampling to 2^7 * 11 / 2^8 * 23
             (1, 2816000), (2, 3328000), (2, 1408000), (4, 1472000)]


@pytest.mark.parametrize("mode,rate", RES_CASES)
def test_resampler_batch(dm, rng, mode, rate):
    """BASELINE configs 3/5 geometry: FIR + resampler, 3 TFs of one stream in one call."""
    bits = bits_for(rng, mode, 3)
    taps = oracle.fir_default_taps()
    ora = oracle.OracleChain(mode=mode, output_rate=rate, fir_taps=taps).run(bits)
    mod = dm.Modulator(mode=mode, output_rate=rate, fir_taps=taps, max_batch=3)
    out = mod.process_batch(bits)
    assert out.shape[1] == ora[0].size
    for i in range(3):
        assert rel_rms(out[i], ora[i]) < TOL, i


@pytest.mark.parametrize("mode,rate", [(1, 8192000), (2, 2500000)])
def test_resampler_state_across_calls(dm, rng, mode, rate):
    """The overlap buffers persist across process() calls (Resampler.cpp:143-145,185-191)."""
    bits = bits_for(rng, mode, 4)
    ora = oracle.OracleChain(mode=mode, output_rate=rate).run(bits)
    mod = dm.Modulator(mode=mode, output_rate=rate, max_batch=2)
    got = [mod.process(bits[0]), mod.process(bits[1])] + list(mod.process_batch(bits[2:]))
    for i in range(4):
        assert rel_rms(got[i], ora[i]) < TOL, i
    # reset() forgets the history: TF 0 again reproduces the first output bit for bit
    mod.reset()
    again = mod.process(bits[0])
    assert np.array_equal(again.view(np.uint32), got[0].view(np.uint32))
    # seek(): a second handle (another GPU in the sharded run) picks the stream up at TF 2
    mod2 = dm.Modulator(mode=mode, output_rate=rate, max_batch=2)
    mod2.seek(2, bits[1])
    tail = mod2.process_batch(bits[2:])
    for i in range(2):
        assert rel_rms(tail[i], ora[2 + i]) < TOL, i


@pytest.mark.parametrize("rate,P", [(4000000, 2), (6000000, 3), (8000000, 4), (10000000, 5)])
def test_resampler_q_kernel(dm, rng, rate, P):
    """k_resample_q (TM I, No = P * 4000: phase transforms of 4000 points in shared memory) is the kernel that runs
    for these rates, agrees with the oracle across calls and after seek(), and with the generic kernel."""
    assert dm.resampler_sizes(2048000, rate, 2048) == (4096, 4000 * P)
    bits = bits_for(rng, 1, 5)
    ora = oracle.OracleChain(mode=1, output_rate=rate).run(bits)
    mod = dm.Modulator(mode=1, output_rate=rate, max_batch=3)
    mod.set_param("profile", 1)
    got = list(mod.process_batch(bits[:3]))
    assert "k_resample_q" in [k for k, _ in mod.kernel_times()]
    got += [mod.process(bits[3]), mod.process(bits[4])]
    for i in range(5):
        assert got[i].size == 196608 * 4000 * P // 4096
        assert rel_rms(got[i], ora[i]) < TOL, i
    mod2 = dm.Modulator(mode=1, output_rate=rate, max_batch=3)
    mod2.seek(3, bits[2])
    for i, y in enumerate(mod2.process_batch(bits[3:])):
        assert np.array_equal(y.view(np.uint32), got[3 + i].view(np.uint32)), i
    gen = dm.Modulator(mode=1, output_rate=rate, max_batch=3)
    gen.set_param("res_kernel", 0)
    gen.set_param("profile", 1)
    ref = gen.process_batch(bits[:3])
    assert "k_resample" in [k for k, _ in gen.kernel_times()]
    for i in range(3):
        assert rel_rms(got[i], ref[i]) < 1e-6, i


def test_resampler_unsupported_rate(dm):
    with pytest.raises(dm.DabModError):
        dm.Modulator(mode=1, output_rate=2048001)


# ---------------------------------------------------------------------------
# MemlessPoly / FormatConverter epilogues, TII, CicEqualizer
# ---------------------------------------------------------------------------
POLY_AM = [1.0, 0.12, -0.3, 0.05, 0.01]
POLY_PM = [0.02, -0.4, 0.3, 0.1, -0.05]


@pytest.mark.parametrize("chain", ["sym", "fir", "res"])
def test_memless_poly(dm, rng, chain):
    kw = dict(mode=2, normalise=1.0 / 46000.0, poly=POLY_AM + POLY_PM)
    if chain != "sym":
        kw["fir_taps"] = oracle.fir_default_taps()
    if chain == "res":
        kw["output_rate"] = 4096000
    bits = bits_for(rng, 2, 2)
    ora = oracle.OracleChain(**kw).run(bits)
    out = dm.Modulator(max_batch=2, **kw).process_batch(bits)
    for i in range(2):
        assert rel_rms(out[i], ora[i]) < TOL


def test_memless_lut(dm, rng):
    lut = (1.0 + 0.1 * rng.standard_normal(32)).astype(np.float32)
    scale = np.float32(2 ** 32 / 1.5)
    kw = dict(mode=2, normalise=1.0 / 46000.0, lut=(scale, lut))
    bits = bits_for(rng, 2, 2)
    ora = oracle.OracleChain(**kw).run(bits)
    out = dm.Modulator(max_batch=2, **kw).process_batch(bits)
    for i in range(2):
        # a magnitude within float rounding of a LUT bin edge may pick the neighbouring entry
        bad = np.abs(out[i] - ora[i]) > 1e-5 * np.abs(ora[i]).max()
        assert bad.sum() <= 4


@pytest.mark.parametrize("fmt,dg", [("s16", 0.8), ("s16", 3.0), ("u8", 0.003), ("s8", 0.003), ("s8", 0.02)])
@pytest.mark.parametrize("chain", ["sym", "fir", "res"])
def test_format_converter(dm, rng, fmt, dg, chain):
    kw = dict(mode=2, digital_gain=dg, fmt=fmt)
    if chain != "sym":
        kw["fir_taps"] = oracle.fir_default_taps()
    if chain == "res":
        kw["output_rate"] = 4096000
    bits = bits_for(rng, 2, 1)
    och = oracle.OracleChain(**kw)
    ora = och.run(bits)[0].astype(np.int32)
    mod = dm.Modulator(**kw)
    out = mod.process_batch(bits)[0].astype(np.int32)
    assert out.size == ora.size
    # float32 inputs differ by ~2e-7 relative: a value next to an integer boundary may truncate to
    # the neighbour.  P(mismatch) = |delta| in LSBs <= 2e-7 * 32767 = 0.0065 per component.
    assert np.abs(out - ora).max() <= 1
    assert np.count_nonzero(out != ora) < 0.01 * ora.size
    # clipped-sample counter (FormatConverter.cpp:176)
    assert abs(int(mod.num_clipped_samples) - int(och.clipped)) <= max(4, 0.002 * och.clipped)


@pytest.mark.parametrize("mode,comb,pattern,old", [(1, 1, 11, 0), (1, 23, 69, 1), (2, 4, 0, 0), (2, 23, 35, 1)])
def test_tii(dm, rng, mode, comb, pattern, old):
    bits = bits_for(rng, mode, 4)
    tii = (comb, pattern, old)
    ora = oracle.OracleChain(mode=mode, tii=tii).run(bits)
    mod = dm.Modulator(mode=mode, tii=tii, max_batch=4)
    out = mod.process_batch(bits)
    null = oracle.mode_params(mode).null_size
    for i in range(4):
        assert rel_rms(out[i], ora[i]) < TOL
        # inserted on every second TF starting with the first (TII.cpp:225-242)
        assert (np.abs(out[i][:null]).max() > 0) == (i % 2 == 0)
    # one TF per call keeps the same parity sequence
    mod.reset()
    for i in range(3):
        assert rel_rms(mod.process(bits[i]), ora[i]) < TOL


def test_tii_modes_without_tii(dm, rng):
    """TM III/IV: the reference's TII constructor throws and NullSymbol is used (DabModulator.cpp:178-190)."""
    bits = bits_for(rng, 4, 1)
    ora = oracle.OracleChain(mode=4).run(bits)
    out = dm.Modulator(mode=4, tii=(1, 1)).process_batch(bits)
    assert rel_rms(out[0], ora[0]) < TOL


@pytest.mark.parametrize("kw", [dict(mode=1, tii=(1, 11, 0), fir_taps="default"), dict(mode=1, tii=(7, 33, 1)),
                                dict(mode=2, tii=(3, 20, 0), fmt="s16", digital_gain=0.8), dict(mode=2, tii=(2, 5, 0), gain_mode="max")],
                         ids=["tm1_fir", "tm1_old_variant", "tm2_s16", "tm2_gain_max"])
def test_tii_on_the_warp_kernels(dm, rng, kw):
    """With 16 or more frames in a call the warp-per-symbol kernels run for TII configurations too: the TII null symbol
    is one constant vector per stream (phase reference carriers x the gain of the phase reference symbol), produced by
    the general kernel on one frame and copied into every second frame (k_tii_fill).  Same frames as the general
    kernel gives for short calls, odd stream offsets included, and as the oracle."""
    mode = kw["mode"]
    n = 17
    bits = bits_for(rng, mode, n)
    fast = dm.Modulator(max_batch=n, **kw)
    fast.set_param("profile", 1)
    a = fast.process_batch(bits)
    names = [k for k, _ in fast.kernel_times()]
    assert "k_tii_fill" in names and ("k_symbols_w" in names or "k_symbols_wg" in names), names
    slow = dm.Modulator(max_batch=n, **kw)
    b = np.concatenate([slow.process_batch(bits[i:i + 5]) for i in range(0, n, 5)])
    assert a.shape == b.shape
    for i in range(n):
        if a.dtype == np.int16:
            d = np.abs(a[i].astype(np.int32) - b[i].astype(np.int32))
            assert d.max() <= 1 and np.count_nonzero(d) < 0.01 * d.size, i
        else:
            assert rel_rms(a[i], b[i]) < 1e-6, i
    # TII frames are the even ones of the stream: after an odd number of frames the next call starts on an odd frame
    c = fast.process_batch(bits)
    slow2 = dm.Modulator(max_batch=n, **kw)
    slow2.seek(n)
    d2 = np.concatenate([slow2.process_batch(bits[i:i + 5]) for i in range(0, n, 5)])
    null = {1: 2656, 2: 664}[mode]
    for i in (0, 1, 2):
        if a.dtype != np.int16:
            assert rel_rms(c[i], d2[i]) < 1e-6, i
        head = (null - 64) * (2 if a.dtype == np.int16 else 1)     # (a FIR reaches 44 samples into symbol 1)
        assert (np.abs(c[i][:head]).max() > 0) == (i % 2 == 1)
    if a.dtype != np.int16:
        okw = dict(kw)
        if okw.get("fir_taps") == "default":
            okw["fir_taps"] = oracle.fir_default_taps()
        want = oracle.OracleChain(**okw).run(bits[:3])
        for i in range(3):
            assert rel_rms(a[i], want[i]) < TOL, i


@pytest.mark.parametrize("clock", [32768000, 400000000, 100000000])
def test_cic_equalizer(dm, rng, clock):
    bits = bits_for(rng, 1, 1)
    ora = oracle.OracleChain(mode=1, clock_rate=clock).run(bits)
    out = dm.Modulator(mode=1, clock_rate=clock).process_batch(bits)
    assert rel_rms(out[0], ora[0]) < TOL


def test_cic_equalizer_fractional_spacing(dm, rng):
    """TM III at 2.5 Msps: N * rate / 2048000 = 312.5, which the reference's `size_t spacing` truncates to 312."""
    bits = bits_for(rng, 3, 1)
    kw = dict(mode=3, clock_rate=100000000, output_rate=2500000)
    ora = oracle.OracleChain(**kw).run(bits)
    out = dm.Modulator(**kw).process_batch(bits)
    assert rel_rms(out[0], ora[0]) < TOL


def test_full_chain_c3_c5(dm, rng):
    """BASELINE configs 3 and 5: FIR + resampler (8.192 / 10 Msps) + MemlessPoly, normalised."""
    bits = bits_for(rng, 1, 2)
    am = [1.0, 0.05, -0.02, 0.0, 0.0]
    pm = [0.0, 0.1, -0.05, 0.0, 0.0]
    for rate in (8192000, 10000000):
        kw = dict(mode=1, output_rate=rate, normalise=1.0 / 46000.0, fir_taps=oracle.fir_default_taps(), poly=am + pm)
        ora = oracle.OracleChain(**kw).run(bits)
        out = dm.Modulator(max_batch=2, **kw).process_batch(bits)
        for i in range(2):
            assert rel_rms(out[i], ora[i]) < TOL


# ---------------------------------------------------------------------------
# Optional features of the same blocks: OFDM windowing, crest factor reduction
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("mode,W", [(1, 10), (1, 504), (2, 7), (3, 4), (3, 63), (4, 32), (4, 252)])
def test_ofdm_windowing(dm, rng, mode, W):
    """GuardIntervalInserter.cpp:149-300; W up to the full guard interval, every chunking."""
    bits = bits_for(rng, mode, 2)
    ora = oracle.OracleChain(mode=mode, window_overlap=W).run(bits)
    mod = dm.Modulator(mode=mode, window_overlap=W, max_batch=2)
    assert mod.get_param("windowlen") == str(W)
    ref = None
    for chunks in (0, 1, 2, 5):
        mod.set_param("sym_chunks", chunks)
        out = mod.process_batch(bits)
        for i in range(2):
            assert rel_rms(out[i], ora[i]) < TOL, (chunks, i)
        if ref is None:
            ref = out.copy()
        else:   # the chunking never changes a bit
            assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), chunks


def test_ofdm_windowing_rc_and_chain(dm, rng):
    """windowlen through the remote-control surface, with TII, FIR and s16 behind it."""
    bits = bits_for(rng, 2, 3)
    taps = oracle.fir_default_taps()
    kw = dict(mode=2, tii=(4, 17, 0), fir_taps=taps, digital_gain=0.8, fmt="s16")
    mod = dm.Modulator(max_batch=3, **kw)
    mod.set_param("windowlen", 20)
    out = mod.process_batch(bits)
    ora = oracle.OracleChain(window_overlap=20, **kw).run(bits)
    for i in range(3):
        d = np.abs(out[i].astype(np.int32) - ora[i].astype(np.int32))
        assert d.max() <= 1 and np.count_nonzero(d) < 1e-3 * d.size, i
    with pytest.raises(dm.DabModError):
        mod.set_param("windowlen", 127)      # TM II guard interval is 126 samples
    mod.set_param("windowlen", 0)
    mod.reset()
    plain = mod.process_batch(bits)
    ora0 = oracle.OracleChain(**kw).run(bits)
    for i in range(3):
        d = np.abs(plain[i].astype(np.int32) - ora0[i].astype(np.int32))
        assert d.max() <= 1, i


@pytest.mark.parametrize("mode,clip,errclip", [(1, 50.0, 0.1), (1, 70.0, 0.02), (2, 25.0, 0.1), (3, 18.0, 0.05),
                                               (4, 35.0, 0.2)])
def test_cfr(dm, rng, mode, clip, errclip):
    """OfdmGeneratorCF32::cfr_one_iteration (OfdmGenerator.cpp:310-373)."""
    bits = bits_for(rng, mode, 2)
    ora = oracle.OracleChain(mode=mode, cfr=(clip, errclip)).run(bits)
    plain = oracle.OracleChain(mode=mode).run(bits)
    assert rel_rms(plain[0], ora[0]) > 1e-3          # the parameters do clip
    mod = dm.Modulator(mode=mode, cfr=(clip, errclip), max_batch=2)
    out = mod.process_batch(bits)
    for i in range(2):
        assert rel_rms(out[i], ora[i]) < TOL, i
    # remote control: switch it off, change the thresholds, switch it on again
    mod.set_param("cfr", 0)
    off = mod.process_batch(bits)
    assert rel_rms(off[0], plain[0]) < TOL
    mod.set_param("clip", clip * 0.8)
    mod.set_param("errorclip", errclip * 2)
    mod.set_param("cfr", 1)
    out2 = mod.process_batch(bits)
    ora2 = oracle.OracleChain(mode=mode, cfr=(clip * 0.8, errclip * 2)).run(bits)
    for i in range(2):
        assert rel_rms(out2[i], ora2[i]) < TOL, i


def test_cfr_with_window_tii_fir_resampler(dm, rng):
    """Everything optional at once on TM I."""
    bits = bits_for(rng, 1, 2)
    taps = oracle.fir_default_taps()
    kw = dict(mode=1, cfr=(50.0, 0.1), window_overlap=16, tii=(1, 11, 0), fir_taps=taps,
              output_rate=4096000, gain_mode="max")
    ora = oracle.OracleChain(**kw).run(bits)
    out = dm.Modulator(max_batch=2, **kw).process_batch(bits)
    for i in range(2):
        assert rel_rms(out[i], ora[i]) < TOL, i


# ---------------------------------------------------------------------------
# TM I has two symbol kernels: warp-per-symbol (default for the plain chain) and
# CTA-per-symbol-group (everything else).  Both must meet the oracle under every chunking.
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("gain_mode", ["var", "max", "fix"])
def test_tm1_symbol_kernels_and_chunking(dm, rng, gain_mode):
    bits = bits_for(rng, 1, 3)
    ora = oracle.OracleChain(mode=1, gain_mode=gain_mode).run(bits)
    mod = dm.Modulator(mode=1, gain_mode=gain_mode, max_batch=3)
    for kernel in (1, 0):
        mod.set_param("sym_kernel", kernel)
        ref = None
        for chunks in (0, 1, 2, 7, 77):
            mod.set_param("sym_chunks", chunks)
            out = mod.process_batch(bits)
            for i in range(3):
                assert rel_rms(out[i], ora[i]) < TOL, (kernel, chunks, i)
            if ref is None:
                ref = out.copy()
            else:
                assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), (kernel, chunks)


def test_tm1_warp_kernel_edge_patterns(dm):
    """All-zero / all-one / alternating bit rows through the warp kernel (phase counters wrap)."""
    m = oracle.mode_params(1)
    pats = np.stack([np.zeros(m.tf_bytes, np.uint8), np.full(m.tf_bytes, 0xff, np.uint8),
                     np.tile(np.array([0xaa, 0x55], np.uint8), m.tf_bytes // 2),
                     np.arange(m.tf_bytes, dtype=np.uint32).astype(np.uint8)])
    ora = oracle.OracleChain(mode=1).run(pats)
    mod = dm.Modulator(mode=1, max_batch=4)
    for chunks in (0, 3):
        mod.set_param("sym_chunks", chunks)
        out = mod.process_batch(pats)
        for i in range(4):
            assert rel_rms(out[i], ora[i]) < TOL, (chunks, i)
