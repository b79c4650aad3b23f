"""TEST INFRASTRUCTURE -- ctypes binding of oracle/liboracle.so (dabmod_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.  The product (odr-dabmod_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

STAGES = {"final": 0, "qpsk": 1, "freq": 2, "diff": 3, "mux": 4, "ciceq": 5, "ofdm": 6,
          "gain": 7, "guard": 8, "fir": 9, "resampler": 10, "poly": 11}
GAIN_MODES = {"fix": 0, "max": 1, "var": 2}
FORMATS = {None: 0, "": 0, "complexf": 0, "s16": 1, "u8": 2, "s8": 3}
FORMAT_DTYPE = {0: np.complex64, 1: np.int16, 2: np.uint8, 3: np.int8}


class Mode(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("mode", "L", "K", "N", "null_size", "sym_size", "beta", "tf_bytes", "tf_samples")]


class Cfg(ctypes.Structure):
    _fields_ = [
        ("mode", ctypes.c_int32),
        ("gain_mode", ctypes.c_int32),
        ("output_rate", ctypes.c_uint64),
        ("clock_rate", ctypes.c_uint64),
        ("digital_gain", ctypes.c_float),
        ("normalise", ctypes.c_float),
        ("gain_variance", ctypes.c_float),
        ("window_overlap", ctypes.c_int32),
        ("cfr_enable", ctypes.c_int32),
        ("cfr_clip", ctypes.c_float),
        ("cfr_errclip", ctypes.c_float),
        ("tii_enable", ctypes.c_int32),
        ("tii_comb", ctypes.c_int32),
        ("tii_pattern", ctypes.c_int32),
        ("tii_old_variant", ctypes.c_int32),
        ("fir_ntaps", ctypes.c_int32),
        ("fir_taps", ctypes.POINTER(ctypes.c_float)),
        ("poly_mode", ctypes.c_int32),
        ("poly_coefs", ctypes.POINTER(ctypes.c_float)),
        ("format", ctypes.c_int32),
        ("fixed_point", ctypes.c_int32),
    ]


def build():
    """Compile liboracle.so (gcc, ~1 s). Building the checker is not using it."""
    srcs = [os.path.join(HERE, f) for f in ("dabmod_oracle.c", "coder_oracle.c", "fixed_oracle.c")]
    if (not os.path.exists(LIB)) or any(os.path.getmtime(LIB) < os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        L.dabo_mode_params.argtypes = [ctypes.c_int, ctypes.POINTER(Mode)]
        L.dabo_chain_new.restype = ctypes.c_void_p
        L.dabo_chain_new.argtypes = [ctypes.POINTER(Cfg)]
        L.dabo_chain_free.argtypes = [ctypes.c_void_p]
        L.dabo_chain_out_samples.restype = ctypes.c_long
        L.dabo_chain_out_samples.argtypes = [ctypes.c_void_p]
        L.dabo_chain_clipped.restype = ctypes.c_uint64
        L.dabo_chain_clipped.argtypes = [ctypes.c_void_p]
        L.dabo_chain_process.restype = ctypes.c_long
        L.dabo_chain_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.dabo_fir_default_taps.argtypes = [ctypes.c_void_p]
        L.dabo_freq_index.argtypes = [ctypes.POINTER(Mode), ctypes.c_void_p]
        L.dabo_phase_ref_index.argtypes = [ctypes.POINTER(Mode), ctypes.c_void_p]
        L.dabo_tii_carriers.argtypes = [ctypes.POINTER(Mode), ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.dabo_dft.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.dabo_cic_filter.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


def mode_params(mode):
    m = Mode()
    if lib().dabo_mode_params(mode, ctypes.byref(m)) != 0:
        raise ValueError("invalid mode %r" % (mode,))
    return m


def fir_default_taps():
    t = np.zeros(45, np.float32)
    lib().dabo_fir_default_taps(t.ctypes.data)
    return t


def freq_index(mode):
    m = mode_params(mode)
    idx = np.zeros(m.K, np.int32)
    lib().dabo_freq_index(ctypes.byref(m), idx.ctypes.data)
    return idx


def phase_ref_index(mode):
    m = mode_params(mode)
    q = np.zeros(m.K, np.uint8)
    lib().dabo_phase_ref_index(ctypes.byref(m), q.ctypes.data)
    return q


def tii_carriers(mode, comb, pattern):
    m = mode_params(mode)
    a = np.zeros(m.K, np.uint8)
    rc = lib().dabo_tii_carriers(ctypes.byref(m), comb, pattern, a.ctypes.data)
    return rc, a


def dft(x, sign):
    x = np.ascontiguousarray(x, np.complex128)
    y = np.empty_like(x)
    lib().dabo_dft(x.size, sign, x.ctypes.data, y.ctypes.data)
    return y


def cic_filter(K, spacing, R):
    f = np.zeros(K, np.float32)
    lib().dabo_cic_filter(K, spacing, R, f.ctypes.data)
    return f


class OracleChain:
    """Stateful (resampler history, TII toggle) oracle instance, one stream."""

    def __init__(self, mode=1, gain_mode="var", output_rate=2048000, clock_rate=0,
                 digital_gain=1.0, normalise=1.0, gain_variance=4.0, window_overlap=0,
                 cfr=None, tii=None, fir_taps=None, poly=None, lut=None, fmt=None, fixed_point=False):
        c = Cfg()
        c.fixed_point = 1 if fixed_point else 0
        c.mode = mode
        c.gain_mode = GAIN_MODES[gain_mode]
        c.output_rate = output_rate
        c.clock_rate = clock_rate
        c.digital_gain = digital_gain
        c.normalise = normalise
        c.gain_variance = gain_variance
        c.window_overlap = window_overlap
        if cfr:
            c.cfr_enable, c.cfr_clip, c.cfr_errclip = 1, cfr[0], cfr[1]
        if tii:
            c.tii_enable, c.tii_comb, c.tii_pattern = 1, tii[0], tii[1]
            c.tii_old_variant = int(tii[2]) if len(tii) > 2 else 0
        self._keep = []
        if fir_taps is not None:
            t = np.ascontiguousarray(fir_taps, np.float32)
            self._keep.append(t)
            c.fir_ntaps = t.size
            c.fir_taps = t.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        if poly is not None:
            p = np.ascontiguousarray(poly, np.float32)
            assert p.size == 10
            self._keep.append(p)
            c.poly_mode = 1
            c.poly_coefs = p.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        elif lut is not None:
            p = np.ascontiguousarray(np.r_[np.float32(lut[0]), np.asarray(lut[1], np.float32)], np.float32)
            assert p.size == 33
            self._keep.append(p)
            c.poly_mode = 2
            c.poly_coefs = p.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        c.format = FORMATS[fmt]
        self.cfg = c
        self.m = mode_params(mode)
        self._h = lib().dabo_chain_new(ctypes.byref(c))
        if not self._h:
            raise RuntimeError("dabo_chain_new failed")
        self.out_samples = lib().dabo_chain_out_samples(self._h)
        self._buf = np.empty(max(self.out_samples, (self.m.L + 1) * self.m.N) * 8 + 64, np.uint8)

    def process(self, bits, stage="final"):
        bits = np.ascontiguousarray(bits, np.uint8)
        assert bits.size == self.m.tf_bytes
        st = STAGES[stage]
        n = lib().dabo_chain_process(self._h, bits.ctypes.data, st, self._buf.ctypes.data)
        if n < 0:
            raise RuntimeError("dabo_chain_process failed")
        dt = np.int16 if self.cfg.fixed_point else FORMAT_DTYPE[self.cfg.format] if st == 0 else np.complex64
        return self._buf[:n].copy().view(dt)

    def run(self, bits_tfs, stage="final"):
        return [self.process(b, stage) for b in np.ascontiguousarray(bits_tfs, np.uint8)]

    @property
    def clipped(self):
        return lib().dabo_chain_clipped(self._h)

    def get_param(self, name):
        """The CFR read-outs "clip_stats" / "papr" (OfdmGenerator.cpp:419-453)."""
        buf = ctypes.create_string_buffer(256)
        L = lib()
        L.dabo_chain_readout.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        if L.dabo_chain_readout(self._h, name.encode(), buf, 256) < 0:
            raise KeyError(name)
        return buf.value.decode()

    def close(self):
        if self._h:
            lib().dabo_chain_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# Channel coding ahead of the path: coder_oracle.c
# ---------------------------------------------------------------------------
MAX_RULES = 8
MAX_STREAMS = 65
ETI_FRAME = 6144


class Rule(ctypes.Structure):
    _fields_ = [("length", ctypes.c_uint32), ("pattern", ctypes.c_uint32)]


class Stream(ctypes.Structure):
    _fields_ = [("framesize", ctypes.c_uint32), ("out_bytes", ctypes.c_uint32), ("start_cu", ctypes.c_uint32),
                ("n_rules", ctypes.c_uint32), ("rules", Rule * MAX_RULES)]

    def as_tuple(self):
        return (self.framesize, self.out_bytes, self.start_cu,
                tuple((self.rules[i].length, self.rules[i].pattern) for i in range(self.n_rules)))


def _coder_lib():
    L = lib()
    if not getattr(L, "_coder_ready", False):
        L.dabc_prbs.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.dabc_conv.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.dabc_puncture.restype = ctypes.c_long
        L.dabc_puncture.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.POINTER(Rule), ctypes.c_int,
                                    ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_long]
        L.dabc_describe.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(Stream), ctypes.c_int]
        L.dabc_coder_new.restype = ctypes.c_void_p
        L.dabc_coder_new.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(Stream)]
        L.dabc_coder_free.argtypes = [ctypes.c_void_p]
        L.dabc_coder_tf_bytes.argtypes = [ctypes.c_void_p]
        L.dabc_coder_feed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L._coder_ready = True
    return L


def prbs(n):
    out = np.zeros(n, np.uint8)
    _coder_lib().dabc_prbs(n, out.ctypes.data)
    return out


def conv_encode(data):
    data = np.ascontiguousarray(data, np.uint8)
    out = np.zeros(4 * data.size + 3, np.uint8)
    _coder_lib().dabc_conv(data.ctypes.data, data.size, out.ctypes.data)
    return out


def puncture(data, rules, out_bytes, tail=(3, 0xcccccc)):
    data = np.ascontiguousarray(data, np.uint8)
    r = (Rule * len(rules))(*[Rule(a, b) for a, b in rules])
    out = np.zeros(out_bytes, np.uint8)
    bits = _coder_lib().dabc_puncture(data.ctypes.data, data.size, r, len(rules), tail[0], tail[1],
                                      out.ctypes.data, out_bytes)
    return out, bits


def describe_eti(frame):
    """(mode, [Stream...]) from the header of one ETI(NI) frame; raises for UEP subchannels."""
    frame = np.ascontiguousarray(frame, np.uint8)
    st = (Stream * MAX_STREAMS)()
    mode = ctypes.c_int()
    n = _coder_lib().dabc_describe(frame.ctypes.data, ctypes.byref(mode), st, MAX_STREAMS)
    if n < 0:
        raise ValueError("dabc_describe: %d (%s)" % (n, "UEP tables are not derived by the oracle" if n == -2 else "bad frame"))
    return mode.value, [st[i] for i in range(n)]


def make_streams(desc):
    """desc: [(framesize, out_bytes, start_cu, ((length, pattern), ...)), ...] -> Stream array"""
    arr = (Stream * len(desc))()
    for i, (fs, ob, sc, rules) in enumerate(desc):
        arr[i].framesize, arr[i].out_bytes, arr[i].start_cu, arr[i].n_rules = fs, ob, sc, len(rules)
        for k, (a, b) in enumerate(rules):
            arr[i].rules[k] = Rule(a, b)
    return arr


class OracleCoder:
    """ETI(NI) frames -> BlockPartitioner blocks, one multiplex configuration."""

    def __init__(self, mode, streams):
        if not isinstance(streams, ctypes.Array):
            streams = make_streams([s.as_tuple() if isinstance(s, Stream) else s for s in streams])
        self._streams = streams
        self._h = _coder_lib().dabc_coder_new(mode, len(streams), streams)
        if not self._h:
            raise ValueError("dabc_coder_new rejected the configuration")
        self.tf_bytes = _coder_lib().dabc_coder_tf_bytes(self._h)
        self._out = np.zeros(self.tf_bytes, np.uint8)

    def feed(self, frame):
        frame = np.ascontiguousarray(frame, np.uint8)
        assert frame.size == ETI_FRAME
        n = _coder_lib().dabc_coder_feed(self._h, frame.ctypes.data, self._out.ctypes.data)
        return self._out.copy() if n else None

    def run(self, frames):
        out = [self.feed(f) for f in np.ascontiguousarray(frames, np.uint8).reshape(-1, ETI_FRAME)]
        return [o for o in out if o is not None]

    def close(self):
        if self._h:
            _coder_lib().dabc_coder_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
