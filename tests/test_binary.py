"""The accelerated engine inside the real program.

oracle/_ref/odr-dabmod-ref  = the complete reference program, every source unmodified (FFTW behind the KISS shim).
oracle/_ref/odr-dabmod-b200 = the same program with `modulator.fft_engine = b200 | b200_fixed`: patched copies of
ConfigParser.cpp / DabModulator.cpp / DabMod.cpp (oracle/patch_engine.py; headers untouched) + the product's adapter.
Both read the same ETI(NI) file through InputFileReader -> EtiReader -> the reference's own channel coding graph and
write an I/Q file through OutputFile; only the chain behind BlockPartitioner differs.  The files must agree TF by TF
(the reference never flushes its P pipelined stages, so its file is P transmission frames shorter).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dabmod_loader  # noqa: E402
from conftest import rel_rms, write_poly_file  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "odr-dabmod-ref")
B200_BIN = os.path.join(ROOT, "oracle", "_ref", "odr-dabmod-b200")
have = os.path.exists(REF_BIN) and os.path.exists(B200_BIN)

INI = """[remotecontrol]
zmqctrl=0
telnet=0
[log]
syslog=0
[input]
transport=file
source={eti}
loop=0
[modulator]
fft_engine={engine}
gainmode={gainmode}
mode={mode}
rate={rate}
digital_gain={digital_gain}
{modulator_extra}
[firfilter]
enabled={fir}
[poly]
enabled={poly}
polycoeffile={polyfile}
num_threads=1
[tii]
enable={tii}
comb=3
pattern=20
[output]
output=file
[fileoutput]
format={fmt}
filename={out}
"""


def eti_mod():
    dabmod_loader.load()
    import importlib
    return importlib.import_module("odr_dabmod_b200.eti")


def run_binary(binary, tmp_path, tag, eti_path, engine, depth=0, **kw):
    cfg = dict(mode=1, rate=2048000, gainmode="var", digital_gain=1.0, fir=0, poly=0, polyfile="/dev/null", tii=0,
               fmt="complexf", modulator_extra="")
    cfg.update(kw)
    out = str(tmp_path / ("out_%s.iq" % tag))
    ini = str(tmp_path / ("cfg_%s.ini" % tag))
    with open(ini, "w") as f:
        f.write(INI.format(eti=eti_path, engine=engine, out=out, **cfg))
    env = dict(os.environ)
    env["ODR_DABMOD_B200_DEPTH"] = str(depth)
    r = subprocess.run([binary, ini], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    return out


def make_eti(tmp_path, mode, n_tf, seed=3):
    e = eti_mod()
    cif = {1: 4, 2: 1, 3: 1, 4: 2}[mode]
    subch = e.default_multiplex() if mode == 1 else [(0, 12, e.eep_tpl(0, 1)), (60, 24, e.eep_tpl(0, 3))]
    frames = e.synth_eti_range(mode, subch, 0, n_tf * cif, seed=seed)
    path = str(tmp_path / "in.eti")
    frames.tofile(path)
    return path


@pytest.mark.skipif(not have, reason="oracle/_ref binaries not built (make -C oracle binary)")
def test_reference_binary_matches_the_reference_harness(tmp_path):
    """CPU: the complete unmodified program and the hot-path harness (oracle/ref_harness.cpp, the checker of every
    parity test) produce the same bytes, so the harness IS the program's chain."""
    from oracle import oracle, refwrap
    if not refwrap.available():
        pytest.skip("reference library not built")
    eti_path = make_eti(tmp_path, 1, 6)
    out = run_binary(REF_BIN, tmp_path, "ref", eti_path, "fftw", fir=1)
    got = np.fromfile(out, np.complex64).reshape(-1, 196608)
    frames = np.fromfile(eti_path, np.uint8).reshape(-1, 6144)
    mode, streams = oracle.describe_eti(frames[0])
    blocks = np.stack(oracle.OracleCoder(mode, [s.as_tuple() for s in streams]).run(frames))
    want = refwrap.RefChain(mode=1, fir_taps_file="default").run(blocks)
    assert got.shape[0] == 6 - 2                      # GainControl and FIRFilter are pipelined: 2 TFs never flushed
    for i in range(got.shape[0]):
        assert np.array_equal(got[i].view(np.uint32), want[i].view(np.uint32)), i


CASES = {
    # name: (mode, n_tf, reference engine, b200 engine, dtype, pipelined stages of the reference, ini settings)
    "c1_native": (1, 6, "fftw", "b200", np.complex64, 1, dict()),
    "c2_fir": (1, 6, "fftw", "b200", np.complex64, 2, dict(fir=1)),
    "c3_fir_res_poly": (1, 5, "fftw", "b200", np.complex64, 3, dict(fir=1, rate=8192000, poly=1, fmt="complexf_normalised")),
    "tm2_s16_tii": (2, 12, "fftw", "b200", np.int16, 1, dict(mode=2, fmt="s16", tii=1, digital_gain=0.8)),
    "fixed_tm1": (1, 5, "kiss", "b200_fixed", np.int16, 0, dict()),
    "fixed_tm4_window": (4, 8, "kiss", "b200_fixed", np.int16, 0, dict(mode=4, modulator_extra="ofdmwindowing=20")),
}


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref binaries not built (make -C oracle binary)")
@pytest.mark.parametrize("case", sorted(CASES))
def test_b200_engine_in_the_real_binary(tmp_path, case):
    mode, n_tf, ref_engine, b200_engine, dt, P, kw = CASES[case]
    if kw.get("poly"):
        kw = dict(kw, polyfile=str(tmp_path / "poly.coef"))
        write_poly_file(kw["polyfile"], [1.0, 0.05, -0.02, 0.0, 0.0], [0.0, 0.1, -0.05, 0.0, 0.0])
    eti_path = make_eti(tmp_path, mode, n_tf)
    ref = np.fromfile(run_binary(REF_BIN, tmp_path, "ref", eti_path, ref_engine, **kw), dt)
    for depth in (0, 3):
        got = np.fromfile(run_binary(B200_BIN, tmp_path, "b200_d%d" % depth, eti_path, b200_engine, depth=depth, **kw), dt)
        assert ref.size % (n_tf - P) == 0
        per_tf = ref.size // (n_tf - P)
        # depth 0: every TF comes back in its own call, nothing is held back; depth D: call i returns TF i - D and
        # the D frames in flight at the end are never flushed, like the reference's pipelined stages
        n_got = n_tf - depth
        assert got.size == n_got * per_tf, (got.size / per_tf, n_got)
        n = min(n_got, n_tf - P)
        assert n >= 2
        a, b = got[:n * per_tf].reshape(n, per_tf), ref[:n * per_tf].reshape(n, per_tf)
        for i in range(n):
            if ref_engine == "kiss":
                assert np.array_equal(a[i], b[i]), (case, depth, i)       # the fixed-point engine is bit-exact
            elif dt is np.int16:
                d = np.abs(a[i].astype(np.int32) - b[i].astype(np.int32))
                assert d.max() <= 1 and np.count_nonzero(d) < 0.01 * d.size, (case, depth, i)
            else:
                assert rel_rms(a[i], b[i]) < 2e-6, (case, depth, i)


ETI_CASES = {
    # name: (case of CASES, engine)
    "c1_native": "b200_eti",
    "c2_fir": "b200_eti",
    "c3_fir_res_poly": "b200_eti",
    "tm2_s16_tii": "b200_eti",
    "fixed_tm1": "b200_eti_fixed",
    "fixed_tm4_window": "b200_eti_fixed",
}


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref binaries not built (make -C oracle binary)")
@pytest.mark.parametrize("case", sorted(ETI_CASES))
def test_eti_engine_in_the_real_binary(tmp_path, case):
    """`fft_engine = b200_eti[_fixed]`: DabModulator's graph is ONE node, B200EtiChain (channel coding + OFDM chain on
    the GPU, fed from the sources EtiReader has parsed), batches of `depth` TFs, the unfinished batch flushed at the
    end of the file.  Every TF of the input comes out; the ones the reference wrote must agree."""
    mode, n_tf, ref_engine, _, dt, P, kw = CASES[case]
    if kw.get("poly"):
        kw = dict(kw, polyfile=str(tmp_path / "poly.coef"))
        write_poly_file(kw["polyfile"], [1.0, 0.05, -0.02, 0.0, 0.0], [0.0, 0.1, -0.05, 0.0, 0.0])
    eti_path = make_eti(tmp_path, mode, n_tf)
    ref = np.fromfile(run_binary(REF_BIN, tmp_path, "ref", eti_path, ref_engine, **kw), dt)
    per_tf = ref.size // (n_tf - P)
    outs = []
    for depth in (2, 4, 64):
        got = np.fromfile(run_binary(B200_BIN, tmp_path, "eti_d%d" % depth, eti_path, ETI_CASES[case], depth=depth, **kw), dt)
        assert got.size == n_tf * per_tf, (got.size / per_tf, n_tf)
        outs.append(got)
        n = n_tf - P
        a, b = got[:n * per_tf].reshape(n, per_tf), ref.reshape(n, per_tf)
        for i in range(n):
            if ref_engine == "kiss":
                assert np.array_equal(a[i], b[i]), (case, depth, i)
            elif dt is np.int16:
                d = np.abs(a[i].astype(np.int32) - b[i].astype(np.int32))
                assert d.max() <= 1 and np.count_nonzero(d) < 0.01 * d.size, (case, depth, i)
            else:
                assert rel_rms(a[i], b[i]) < 2e-6, (case, depth, i)
    # the batch size changes nothing
    # (gain mode var: per-symbol statistics; the resampler state is carried across batches)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2]), case


B1_BIN = os.path.join(ROOT, "oracle", "_ref", "odr-dabmod-b1")


@pytest.mark.gpu
@pytest.mark.skipif(not (have and os.path.exists(B1_BIN)), reason="oracle/_ref binaries not built (make -C oracle binary b1)")
@pytest.mark.parametrize("case", sorted(CASES))
def test_translation_unit_substitution(tmp_path, case):
    """Boundary shape B1: oracle/_ref/odr-dabmod-b1 is the reference program with DabModulator.cpp,
    ConfigParser.cpp and DabMod.cpp UNMODIFIED; only the sixteen translation units of the hot path are replaced by
    adapter/B200Blocks.cpp (same classes, reference headers, the chain on the GPU through the C ABI).  Same
    configuration file as the reference (fft_engine = fftw | kiss), same remote-control surface, same pipeline
    delays: the two output files have the same length and agree frame by frame."""
    mode, n_tf, ref_engine, _, dt, P, kw = CASES[case]
    if kw.get("poly"):
        kw = dict(kw, polyfile=str(tmp_path / "poly.coef"))
        write_poly_file(kw["polyfile"], [1.0, 0.05, -0.02, 0.0, 0.0], [0.0, 0.1, -0.05, 0.0, 0.0])
    eti_path = make_eti(tmp_path, mode, n_tf)
    ref = np.fromfile(run_binary(REF_BIN, tmp_path, "ref", eti_path, ref_engine, **kw), dt)
    got = np.fromfile(run_binary(B1_BIN, tmp_path, "b1", eti_path, ref_engine, **kw), dt)
    assert got.size == ref.size, (got.size, ref.size)
    n = n_tf - P
    a, b = got.reshape(n, -1), ref.reshape(n, -1)
    for i in range(n):
        if ref_engine == "kiss":
            assert np.array_equal(a[i], b[i]), (case, i)
        elif dt is np.int16:
            d = np.abs(a[i].astype(np.int32) - b[i].astype(np.int32))
            assert d.max() <= 1 and np.count_nonzero(d) < 0.01 * d.size, (case, i)
        else:
            assert rel_rms(a[i], b[i]) < 2e-6, (case, i)


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not (have and os.path.exists(B1_BIN)), reason="oracle/_ref binaries not built")
@pytest.mark.skipif(not _no_gpu(), reason="needs a host without a GPU")
@pytest.mark.parametrize("binary,engine", [(B1_BIN, "fftw"), (B200_BIN, "b200"), (B200_BIN, "b200_eti")])
def test_binaries_fail_loudly_without_a_gpu(tmp_path, binary, engine):
    """CPU: there is no fallback behind any of the three bindings -- the program stops with the library's error
    instead of computing on the host."""
    eti_path = make_eti(tmp_path, 1, 4)
    out = str(tmp_path / "out.iq")
    ini = str(tmp_path / "cfg.ini")
    cfg = dict(mode=1, rate=2048000, gainmode="var", digital_gain=1.0, fir=0, poly=0, polyfile="/dev/null", tii=0,
               fmt="complexf", modulator_extra="")
    with open(ini, "w") as f:
        f.write(INI.format(eti=eti_path, engine=engine, out=out, **cfg))
    r = subprocess.run([binary, ini], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    # (the reference's logger runs on its own thread and does not always get the library's message out before exit)
    assert r.returncode != 0, r.stdout[-1500:]
    assert not os.path.exists(out) or os.path.getsize(out) == 0


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref binaries not built (make -C oracle binary)")
def test_eti_engine_survives_a_multiplex_reconfiguration(tmp_path):
    """The ensemble changes in mid-stream (six subchannels -> two).  The reference notices in FrameMultiplexer, restarts
    the modulator and resumes at the next frame with FP 0 (src/DabMod.cpp:575-579, 691-700, 744-749).  The ETI engine
    does the same after draining what it has collected: its output is the first part complete, then the second part
    from the frame the new modulator starts on -- each exactly what a run on that part alone produces."""
    e = eti_mod()
    n1, n2 = 40, 64                                         # ETI frames (TM I: 10 and 16 TFs); n1 a multiple of 8
    part1 = e.synth_eti_range(1, e.default_multiplex(), 0, n1, seed=5)
    part2 = e.synth_eti_range(1, [(0, 12, e.eep_tpl(0, 1)), (60, 24, e.eep_tpl(0, 3))], n1, n2, seed=6)
    paths = {}
    for name, frames in (("both", np.concatenate([part1, part2])), ("p1", part1), ("p2", part2[8:])):
        paths[name] = str(tmp_path / (name + ".eti"))
        frames.tofile(paths[name])
    out = {k: np.fromfile(run_binary(B200_BIN, tmp_path, "eti_" + k, v, "b200_eti", depth=4), np.complex64)
           for k, v in paths.items()}
    per_tf = 196608
    assert out["p1"].size == (n1 // 4) * per_tf and out["p2"].size == ((n2 - 8) // 4) * per_tf
    assert out["both"].size == out["p1"].size + out["p2"].size
    assert np.array_equal(out["both"][:out["p1"].size].view(np.uint32), out["p1"].view(np.uint32))
    assert np.array_equal(out["both"][out["p1"].size:].view(np.uint32), out["p2"].view(np.uint32))
    # and the reference program on the same file: same frames (it never flushes its one pipelined TF per modulator)
    ref = np.fromfile(run_binary(REF_BIN, tmp_path, "ref_both", paths["both"], "fftw"), np.complex64)
    r1 = n1 // 4 - 1
    assert ref.size == (r1 + (n2 - 8) // 4 - 1) * per_tf
    for i in range(r1):
        assert rel_rms(out["both"][i * per_tf:(i + 1) * per_tf], ref[i * per_tf:(i + 1) * per_tf]) < 2e-6, i
    for i in range((n2 - 8) // 4 - 1):
        a = out["both"][out["p1"].size + i * per_tf:out["p1"].size + (i + 1) * per_tf]
        assert rel_rms(a, ref[(r1 + i) * per_tf:(r1 + i + 1) * per_tf]) < 2e-6, i


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree is only present in the build container")
def test_patch_engine_anchors_still_match(tmp_path):
    """CPU (build container): oracle/patch_engine.py finds every anchor in the reference sources exactly once and
    writes the three patched copies with both engine families in them."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "patch_engine.py"), "/root/reference", str(tmp_path)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    cfg = open(tmp_path / "ConfigParser.cpp").read()
    mod = open(tmp_path / "DabModulator.cpp").read()
    main = open(tmp_path / "DabMod.cpp").read()
    for name in ("b200", "b200_fixed", "b200_eti", "b200_eti_fixed"):
        assert '"%s"' % name in cfg
    assert "B200OfdmChain" in mod and "B200EtiChain" in mod and "B200SwapOutput" in mod
    assert "B200EtiChain::flush_active()" in main
    # nothing but the anchored edits: the copies differ from the reference in a handful of lines
    import difflib
    for name, text in (("ConfigParser.cpp", cfg), ("DabModulator.cpp", mod), ("DabMod.cpp", main)):
        ref = open(os.path.join("/root/reference/src", name)).read()
        changed = [l for l in difflib.unified_diff(ref.splitlines(), text.splitlines(), lineterm="", n=0)
                   if l[:1] in "+-" and l[:3] not in ("+++", "---")]
        assert 0 < len(changed) < 80, (name, len(changed))
