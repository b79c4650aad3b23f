"""Channel coding ahead of the hot path.

CPU: the oracle restatement (oracle/coder_oracle.c) is pinned bit for bit against the
unmodified reference (EtiReader + PrbsGenerator/ConvEncoder/PuncturingEncoder/
TimeInterleaver/FrameMultiplexer/BlockPartitioner in the reference's Flowgraph,
oracle/ref_coder_harness.cpp), and against the committed golden blocks.
GPU: the CUDA coder through the C ABI is compared bit for bit with the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dabmod_loader  # noqa: E402
from oracle import oracle, refwrap  # noqa: E402

have_ref = refwrap.available()
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coder_blocks.npz")


def eti_mod():
    dabmod_loader.load()
    import importlib
    return importlib.import_module("odr_dabmod_b200.eti")


# (mode, subchannels [(sad, stl, tpl)]) -- EEP-A levels 1-4 at several rates, EEP-B, gaps, overlap order
def multiplexes():
    e = eti_mod()
    return {
        "tm1_six_128k_3a": (1, e.default_multiplex()),
        "tm2_mixed_eep_a": (2, [(0, 12, e.eep_tpl(0, 1)), (60, 3, e.eep_tpl(0, 2)), (70, 24, e.eep_tpl(0, 2)),
                                (200, 36, e.eep_tpl(0, 4)), (400, 72, e.eep_tpl(0, 3))]),
        "tm3_eep_b": (3, [(0, 12, e.eep_tpl(1, 1)), (30, 24, e.eep_tpl(1, 2)), (100, 48, e.eep_tpl(1, 3)),
                          (300, 96, e.eep_tpl(1, 4))]),
        "tm4_fic_only": (4, []),
        "tm1_one_big": (1, [(0, 216, e.eep_tpl(0, 3))]),        # 576 kbit/s, 432 CU
    }


UEP_MUX = (1, [(0, 12, 0x02), (40, 48, 0x12), (200, 72, 0x0b)])   # short form (UEP) subchannels


@pytest.mark.parametrize("n,inv", [(1, 0), (96, 0), (6912, 0)])
def test_prbs_properties(n, inv):
    p = oracle.prbs(n)
    # x^9 + x^5 + 1 from the all-ones state (EN 300 401 10.1): first bits 0000 0111 1011 1110 ...
    if n >= 2:
        assert p[0] == 0x07 and p[1] == 0xBE
    bits = np.unpackbits(oracle.prbs(128))
    assert np.array_equal(bits[:511], bits[511:1022])            # period 511


def test_conv_encoder_known_answers():
    # all-zero input -> all-zero output; a single one -> the generator taps in turn
    assert not oracle.conv_encode(np.zeros(4, np.uint8)).any()
    out = np.unpackbits(oracle.conv_encode(np.array([0x80, 0, 0, 0], np.uint8)))
    want = []
    for t in range(7):                     # the one moves from register bit 6 down to bit 0
        reg = 0x40 >> t
        want += [bin(reg & g).count("1") & 1 for g in (0x5b, 0x79, 0x65, 0x5b)]
    assert out[:28].tolist() == want and not out[28:].any()


@pytest.mark.skipif(not have_ref, reason="reference library not built")
@pytest.mark.parametrize("name", sorted(multiplexes()))
def test_oracle_matches_reference(name):
    mode, subch = multiplexes()[name]
    cif = {1: 4, 2: 1, 3: 1, 4: 2}[mode]
    frames = eti_mod().synth_eti(mode, subch, 20 * cif if name != "tm1_one_big" else 24, seed=hash(name) & 0xffff)
    ref = refwrap.RefCoder()
    want = ref.run(frames)
    m2, streams = oracle.describe_eti(frames[0])
    assert m2 == mode
    assert [s.as_tuple() for s in streams] == ref.describe()      # sizes and puncturing rules
    got = oracle.OracleCoder(mode, streams).run(frames)
    assert len(got) == len(want) == frames.shape[0] // cif
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), (name, i, int(np.argmax(a != b)))


@pytest.mark.skipif(not have_ref, reason="reference library not built")
def test_oracle_matches_reference_uep_rules_from_reference():
    """UEP subchannels: the rules come from the reference's SubchannelSource (as the C++ adapter passes them)."""
    mode, subch = UEP_MUX
    frames = eti_mod().synth_eti(mode, subch, 80, seed=77)
    ref = refwrap.RefCoder()
    want = ref.run(frames)
    with pytest.raises(ValueError):
        oracle.describe_eti(frames[0])
    got = oracle.OracleCoder(mode, ref.describe()).run(frames)
    assert len(got) == len(want) == 20
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_oracle_matches_golden():
    g = np.load(GOLDEN)
    for name in sorted(multiplexes()):
        mode, subch = multiplexes()[name]
        frames = eti_mod().synth_eti(mode, subch, int(g[name + "/n_frames"]), seed=int(g[name + "/seed"]))
        _, streams = oracle.describe_eti(frames[0])
        got = np.stack(oracle.OracleCoder(mode, streams).run(frames))
        assert got.shape == tuple(g[name + "/shape"])
        assert np.array_equal(got[-2:], g[name + "/last2"])
        assert np.array_equal(got.astype(np.uint64).sum(axis=1), g[name + "/rowsum"])


# ---------------------------------------------------------------------------
# Product side.  Host-only pieces (header parsing, rule derivation) on the CPU ...
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(multiplexes()))
def test_eti_describe_matches_oracle(name):
    dm = dabmod_loader.load()
    mode, subch = multiplexes()[name]
    frame = eti_mod().synth_eti(mode, subch, 1)[0]
    m_o, st_o = oracle.describe_eti(frame)
    m_p, st_p = dm.eti_describe(frame)
    assert m_p == m_o == mode
    assert st_p == [s.as_tuple() for s in st_o]


def test_eti_describe_errors():
    dm = dabmod_loader.load()
    frame = eti_mod().synth_eti(*UEP_MUX, 1)[0]
    with pytest.raises(dm.DabModError) as e:
        dm.eti_describe(frame)
    assert e.value.code == -4                       # UEP tables: rules must be passed explicitly
    bad = frame.copy()
    bad[5] &= 0x7f                                  # FICF = 0 (EtiReader.cpp:143-145)
    with pytest.raises(dm.DabModError):
        dm.eti_describe(bad)


# ... and the kernels on the GPU, bit for bit against the oracle
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(multiplexes()))
def test_cuda_coder_matches_oracle(name):
    dm = dabmod_loader.load()
    mode, subch = multiplexes()[name]
    cif = {1: 4, 2: 1, 3: 1, 4: 2}[mode]
    frames = eti_mod().synth_eti(mode, subch, 24 * cif, seed=5)
    _, streams = dm.eti_describe(frames[0])
    want = np.stack(oracle.OracleCoder(mode, streams).run(frames))
    # one call, then the same stream in uneven pieces (history carried across calls)
    cod = dm.Coder(mode, streams, max_frames=24 * cif)
    got = cod.process(frames)
    assert np.array_equal(got, want)
    cod.reset()
    pieces, pos = [], 0
    for n_tf in (1, 3, 2, 7, 11):
        pieces.append(cod.process(frames[pos:pos + n_tf * cif]))
        pos += n_tf * cif
    assert pos == frames.shape[0]
    assert np.array_equal(np.concatenate(pieces), want)
    # golden blocks of the unmodified reference
    g = np.load(GOLDEN)
    gf = eti_mod().synth_eti(mode, subch, int(g[name + "/n_frames"]), seed=int(g[name + "/seed"]))
    cod.reset()
    gb = cod.process(gf)
    assert np.array_equal(gb[-2:], g[name + "/last2"])
    assert np.array_equal(gb.astype(np.uint64).sum(axis=1), g[name + "/rowsum"])


@pytest.mark.gpu
def test_cuda_coder_uep_rules_and_errors():
    dm = dabmod_loader.load()
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "coder_uep.npz"))
    streams = [(int(a), int(b), int(c), tuple((int(x), int(y)) for x, y in r.reshape(-1, 2)))
               for a, b, c, r in zip(g["framesize"], g["out_bytes"], g["start_cu"],
                                     [g["rules%d" % i] for i in range(len(g["framesize"]))])]
    frames = eti_mod().synth_eti(*UEP_MUX, 80, seed=77)
    cod = dm.Coder(UEP_MUX[0], streams, max_frames=80)
    got = cod.process(frames)
    assert np.array_equal(got, np.stack(oracle.OracleCoder(UEP_MUX[0], streams).run(frames)))
    assert np.array_equal(got[-2:], g["last2"])
    with pytest.raises(dm.DabModError):
        cod.process(frames[:3])                      # not a whole transmission frame
    with pytest.raises(dm.DabModError):
        dm.Coder(1, [(96, 200, 0, ((21 * 16, 0xeeeeeeee), (3 * 16, 0xeeeeeeec)))])   # FIC size mismatch


@pytest.mark.gpu
def test_cuda_coder_shard_priming():
    """A shard of the stream: prime the time interleaver with the 15 frames before it."""
    dm = dabmod_loader.load()
    mode, subch = multiplexes()["tm1_six_128k_3a"]
    frames = eti_mod().synth_eti(mode, subch, 96, seed=9)
    _, streams = dm.eti_describe(frames[0])
    want = np.stack(oracle.OracleCoder(mode, streams).run(frames))
    cod = dm.Coder(mode, streams, max_frames=48)
    cod.prime(frames[:48])                           # more than needed: only the last 15 count
    assert np.array_equal(cod.process(frames[48:]), want[12:])
    cod.prime(frames[40 - 15:40])
    assert np.array_equal(cod.process(frames[40:88]), want[10:22])


@pytest.mark.gpu
def test_eti_to_iq_on_device():
    """ETI bytes -> I/Q with the coded blocks staying on the device, vs oracle coder + oracle chain."""
    from conftest import rel_rms
    dm = dabmod_loader.load()
    mode, subch = multiplexes()["tm1_six_128k_3a"]
    frames = eti_mod().synth_eti(mode, subch, 16, seed=3)
    _, streams = dm.eti_describe(frames[0])
    blocks = oracle.OracleCoder(mode, streams).run(frames)
    taps = oracle.fir_default_taps()
    want = oracle.OracleChain(mode=mode, fir_taps=taps).run(np.stack(blocks))
    mod = dm.Modulator(mode=mode, fir_taps=taps, max_batch=4)
    cod = dm.Coder(mode, streams, max_frames=16)
    got = cod.modulate(mod, frames)
    assert got.shape[0] == 4
    for i in range(4):
        assert rel_rms(got[i], want[i]) < 2e-6, i


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tm1_six_128k_3a", "tm2_mixed_eep_a", "tm3_eep_b"])
def test_eti_to_fixed_point_iq_on_device(name):
    """The whole reference graph for FFTEngine::KISS -- ETI bytes -> channel coding -> fixed-point OFDM -- on the
    device: every bit of the int16 I/Q equals oracle coder + fixed-point oracle (both pinned against the reference)."""
    dm = dabmod_loader.load()
    mux = multiplexes()
    if name not in mux:
        name = sorted(mux)[0]
    mode, subch = mux[name]
    per_tf = {1: 4, 2: 1, 3: 1, 4: 2}[mode]
    frames = eti_mod().synth_eti(mode, subch, 4 * per_tf, seed=5)
    _, streams = dm.eti_describe(frames[0])
    blocks = oracle.OracleCoder(mode, streams).run(frames)
    want = oracle.OracleChain(mode=mode, fixed_point=True, window_overlap=8).run(np.stack(blocks))
    mod = dm.Modulator(mode=mode, fixed_point=True, window_overlap=8, max_batch=4)
    cod = dm.Coder(mode, streams, max_frames=4 * per_tf)
    got = cod.modulate(mod, frames)
    assert got.dtype == np.int16 and got.shape[0] == 4
    for i in range(4):
        assert np.array_equal(got[i], want[i]), i
