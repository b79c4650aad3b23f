"""Row N4: the fixed-point engine (FFTEngine::KISS, DabModulator.cpp:144-224) on the GPU.

Integer arithmetic from the bits to the int16 I/Q: the CUDA path (k_symbols_fix through the C ABI) must equal
the oracle (oracle/fixed_oracle.c, pinned bit-exact against the compiled reference in
tests/test_oracle_vs_reference.py and against tests/golden/fix_*.npz) in every bit."""
import numpy as np
import pytest

import dabmod_loader
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dm():
    return dabmod_loader.load()


def bits_for(rng, mode, n):
    return rng.integers(0, 256, (n, oracle.mode_params(mode).tf_bytes), dtype=np.uint8)


CASES = [dict(), dict(window_overlap=16), dict(tii=(3, 5, 0)), dict(window_overlap=7, tii=(20, 33, 1))]


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("kw", CASES, ids=["plain", "window", "tii", "window_tii_old"])
def test_fixed_point_bit_exact(dm, rng, mode, kw):
    if mode > 2:
        kw = {k: v for k, v in kw.items() if k != "tii"}       # TM III / IV: TIIError -> plain null symbol
    bits = bits_for(rng, mode, 5)
    want = oracle.OracleChain(mode=mode, fixed_point=True, **kw).run(bits)
    mod = dm.Modulator(mode=mode, fixed_point=True, max_batch=5, **kw)
    assert mod.tf_out_bytes == 4 * oracle.mode_params(mode).tf_samples
    got = mod.process_batch(bits)
    assert got.dtype == np.int16
    for i in range(5):
        assert np.array_equal(got[i], want[i]), i


@pytest.mark.parametrize("mode,W", [(1, 0), (1, 504), (2, 126), (3, 63), (4, 252)])
def test_fixed_point_chunking_and_calls(dm, rng, mode, W):
    """The work split (CTAs per TF), the call granularity and the device-buffer entry point never change a bit;
    W up to the whole guard interval."""
    bits = bits_for(rng, mode, 4)
    want = oracle.OracleChain(mode=mode, fixed_point=True, window_overlap=W, tii=(1, 2, 0) if mode < 3 else None).run(bits)
    mod = dm.Modulator(mode=mode, fixed_point=True, window_overlap=W, tii=(1, 2, 0) if mode < 3 else None, max_batch=4)
    for chunks in (0, 1, 3, 11):
        mod.reset()
        mod.set_param("sym_chunks", chunks)
        got = mod.process_batch(bits)
        for i in range(4):
            assert np.array_equal(got[i], want[i]), (chunks, i)
    mod.reset()
    singles = [mod.process(b) for b in bits]          # TII parity advances call by call
    for i in range(4):
        assert np.array_equal(singles[i], want[i]), i
    import torch
    mod.reset()
    d_in = torch.from_numpy(bits).cuda()
    d_out = torch.empty(4 * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
    mod.process_batch_device(d_in.data_ptr(), 4, d_out.data_ptr())
    mod.synchronize()
    dev = d_out.cpu().numpy().view(np.int16).reshape(4, -1)
    for i in range(4):
        assert np.array_equal(dev[i], want[i]), i


def test_fixed_point_windowlen_rc(dm, rng):
    """windowlen is remote-controllable in the fixed chain too (GuardIntervalInserter.cpp:338-379)."""
    bits = bits_for(rng, 2, 2)
    mod = dm.Modulator(mode=2, fixed_point=True, max_batch=2)
    a = mod.process_batch(bits)
    mod.set_param("windowlen", 24)
    b = mod.process_batch(bits)
    want0 = oracle.OracleChain(mode=2, fixed_point=True).run(bits)
    want1 = oracle.OracleChain(mode=2, fixed_point=True, window_overlap=24).run(bits)
    for i in range(2):
        assert np.array_equal(a[i], want0[i]) and np.array_equal(b[i], want1[i])
    assert mod.get_param("windowlen") == "24"


def test_fixed_point_rejections(dm):
    """DabModulator.cpp:249,257,265: "fixed point doesn't support ..."; no GainControl / CFR in that chain."""
    for kw in (dict(fir_taps="default"), dict(output_rate=4096000), dict(poly=[1, 0, 0, 0, 0, 0, 0, 0, 0, 0]),
               dict(cfr=(50.0, 0.1)), dict(clock_rate=32768000), dict(fmt="u8")):
        with pytest.raises(dm.DabModError):
            dm.Modulator(mode=1, fixed_point=True, **kw)
    mod = dm.Modulator(mode=1, fixed_point=True)
    for name, value in (("cfr", "1"), ("digital", "0.5"), ("mode", "max"), ("taps", "1 1.0")):
        with pytest.raises(dm.DabModError):
            mod.set_param(name, value)
    with pytest.raises(dm.DabModError):
        mod.process(np.zeros(17, np.uint8))
