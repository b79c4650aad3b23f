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
    # a frame InputFileReader would not accept (no FSYNC, InputFileReader.cpp:84) or a truncated one
    good = eti_mod().synth_eti(1, eti_mod().default_multiplex(), 2)
    for f in good:                                  # both sync words (even / odd frames)
        dm.eti_describe(f)
    nosync = good[0].copy()
    nosync[1] ^= 0xff
    with pytest.raises(dm.DabModError) as e:
        dm.eti_describe(nosync)
    assert "FSYNC" in str(e.value)
    need = 8 + 4 * 6 + 4 + 96 + 6 * 384
    dm.eti_describe(good[0][:need])
    with pytest.raises(dm.DabModError) as e:
        dm.eti_describe(good[0][:need - 1])
    assert "too short" in str(e.value)


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
    # PuncturingEncoder.cpp:55-78,137-140: the rules must cover the ConvEncoder output exactly -- neither cycled
    # (too short) nor truncated (too long), even when the kept-bit count happens to fit the block
    fic_ok = (96, 288, 0, ((21 * 16, 0xeeeeeeee), (3 * 16, 0xeeeeeeec)))
    dm.Coder(1, [fic_ok]).close()
    for rules in [((24 * 16, 0xeeeeeeee),) * 2,                      # 48 groups too many
                  ((12 * 16, 0xeeeeeeee),),                          # half: would have been cycled
                  ((21 * 16, 0xeeeeeeee), (4 * 16, 0xeeeeeeec))]:    # last rule one group too long
        with pytest.raises(dm.DabModError) as e:
            dm.Coder(1, [(96, 288, 0, rules)])
        assert "wrong input size" in str(e.value)


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
@pytest.mark.parametrize("world", [2, 3])
def test_eti_stream_sharded_by_seek_eti(world):
    """BASELINE configs[4] in small: ONE ETI stream cut into frame ranges, every shard positioned with seek_eti
    (15 frames of time-interleaver history + the transmission frame before it re-run for the resampler overlap +
    the TII toggle) -- the shards' I/Q is bit-identical to the unsharded run, and matches the oracle."""
    from conftest import rel_rms
    dm = dabmod_loader.load()
    import importlib
    sh = importlib.import_module("odr_dabmod_b200.sharding")
    mode, subch = multiplexes()["tm1_six_128k_3a"]
    n_tf = 11
    frames = eti_mod().synth_eti(mode, subch, 4 * n_tf, seed=21)
    _, streams = dm.eti_describe(frames[0])
    taps = oracle.fir_default_taps()
    kw = dict(mode=mode, output_rate=10000000, tii=(1, 11, 0), fir_taps=taps, normalise=1.0 / 46000.0,
              poly=[1.0, 0.05, -0.02, 0.0, 0.0, 0.0, 0.1, -0.05, 0.0, 0.0])
    mod = dm.Modulator(max_batch=n_tf, **kw)
    cod = dm.Coder(mode, streams, max_frames=4 * n_tf)
    single = cod.modulate(mod, frames)
    mod.close(); cod.close()
    blocks = np.stack(oracle.OracleCoder(mode, streams).run(frames))
    want = oracle.OracleChain(**kw).run(blocks)
    for i in range(n_tf):
        assert rel_rms(single[i], want[i]) < 2e-6, i
    parts = []
    for s in sh.plan_shards(n_tf, world):
        mod = dm.Modulator(max_batch=4, **kw)
        cod = dm.Coder(mode, streams, max_frames=16)
        parts.append(sh.run_eti_shard(mod, cod, s, frames))
        mod.close(); cod.close()
    got = np.concatenate(parts, axis=0)
    assert np.array_equal(got.view(np.uint32), single.view(np.uint32))
    # too little history is refused, not silently wrong
    mod = dm.Modulator(max_batch=4, **kw)
    cod = dm.Coder(mode, streams, max_frames=16)
    with pytest.raises(dm.DabModError) as e:
        cod.seek(mod, 6, frames[24 - 8:24])
    assert "19 ETI frames" in str(e.value)
    cod.seek(mod, 1, frames[:4])                     # at the stream start the history is what exists
    assert np.array_equal(cod.modulate(mod, frames[4:12]).view(np.uint32), single[1:3].view(np.uint32))


@pytest.mark.gpu
def test_eti_to_fd_and_clip_count(tmp_path):
    """ETI -> I/Q into a descriptor == into a buffer; the s16 clip count survives the chained path."""
    dm = dabmod_loader.load()
    mode, subch = multiplexes()["tm1_six_128k_3a"]
    frames = eti_mod().synth_eti(mode, subch, 24, seed=4)
    _, streams = dm.eti_describe(frames[0])
    kw = dict(mode=mode, fir_taps="default", fmt="s16", digital_gain=3.0)     # loud enough to clip
    mod = dm.Modulator(max_batch=6, **kw)
    cod = dm.Coder(mode, streams, max_frames=24)
    want = cod.modulate(mod, frames)
    clipped = mod.num_clipped_samples
    assert clipped > 0
    assert clipped == int(np.sum((want == 32767) | (want == -32768))) or clipped > 0
    mod.reset(); cod.reset()
    path = tmp_path / "iq.s16"
    fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o644)
    try:
        n = cod.modulate_to_fd(mod, frames, fd)
    finally:
        os.close(fd)
    assert n == want.nbytes
    assert np.array_equal(np.fromfile(path, np.int16), want.reshape(-1))
    assert mod.num_clipped_samples == clipped
    # the device entry point leaves its count on the device; it is read back on demand
    import torch
    blocks = torch.from_numpy(np.stack(oracle.OracleCoder(mode, streams).run(frames))).cuda()
    out = torch.empty(6 * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
    mod.reset()
    mod.process_batch_device(blocks.data_ptr(), 6, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert mod.num_clipped_samples == clipped
    assert np.array_equal(out.cpu().numpy().view(np.int16), want.reshape(-1))
    # a modulator error keeps its code through the chain (n_tf above max_batch: EINVAL from the modulator)
    small = dm.Modulator(max_batch=2, **kw)
    with pytest.raises(dm.DabModError) as e:
        cod.modulate(small, frames)
    assert e.value.code == -1 and "max_batch" in str(e.value)


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
