"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_ref/libdabmod_ref.so.

That library is the UNMODIFIED reference hot-path code (compiled from
/root/reference by oracle/Makefile) driven through the reference's own
Flowgraph by oracle/ref_harness.cpp.  It exists to pin our C restatement
(oracle/dabmod_oracle.c), to generate the golden vectors in tests/golden/ and
as the CPU baseline of bench.py.  Nothing in the product imports this.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libdabmod_ref.so")
REF_LIB_FAST = os.path.join(HERE, "_ref", "fast", "libdabmod_ref.so")     # the same sources with -ffast-math
# the same harness linked with the product's adapter/B200Blocks.cpp in place of the sixteen hot-path translation units
# (boundary shape B1): needs a GPU, used by tests/test_b1_blocks.py as the thing under test, never as a checker
B1_LIB = os.path.join(HERE, "_ref", "libdabmod_b1.so")

TF_BYTES = {1: 75 * 384, 2: 75 * 96, 3: 152 * 48, 4: 75 * 192}
TF_SAMPLES = {1: 196608, 2: 49152, 3: 49152, 4: 98304}


class _RefCfg(ctypes.Structure):
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
        ("poly_threads", ctypes.c_int32),
        ("fir_taps_file", ctypes.c_char_p),
        ("poly_coef_file", ctypes.c_char_p),
        ("format", ctypes.c_char_p),
        ("stop_after", ctypes.c_char_p),
        ("fixed_point", ctypes.c_int32),
    ]


def available(variant=None):
    return os.path.exists({"fast": REF_LIB_FAST, "b1": B1_LIB}.get(variant, REF_LIB))


_lib = None
_libs = {}


def lib(variant=None):
    global _lib
    if variant in ("fast", "b1"):
        if variant not in _libs:
            saved, _lib = _lib, None
            try:
                _libs[variant] = _load(REF_LIB_FAST if variant == "fast" else B1_LIB)
            finally:
                _lib = saved
        return _libs[variant]
    if _lib is None:
        _lib = _load(REF_LIB)
    return _lib


def _load(path):
    if True:
        _lib = ctypes.CDLL(path)
        _lib.ref_create.restype = ctypes.c_void_p
        _lib.ref_create.argtypes = [ctypes.POINTER(_RefCfg)]
        _lib.ref_latency.restype = ctypes.c_int
        _lib.ref_latency.argtypes = [ctypes.c_void_p]
        _lib.ref_process.restype = ctypes.c_long
        _lib.ref_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_void_p, ctypes.c_size_t]
        _lib.ref_destroy.argtypes = [ctypes.c_void_p]
        _lib.ref_last_error.restype = ctypes.c_char_p
        _lib.ref_ofdm_get_parameter.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    return _lib


GAIN_MODES = {"fix": 0, "max": 1, "var": 2}


class RefChain:
    """One reference flowgraph instance (stateful: resampler history, TII toggle)."""

    def __init__(self, mode=1, gain_mode="var", output_rate=2048000, clock_rate=0,
                 digital_gain=1.0, normalise=1.0, gain_variance=4.0, window_overlap=0,
                 cfr=None, tii=None, fir_taps_file=None, poly_coef_file=None,
                 poly_threads=1, fmt=None, stop_after=None, fixed_point=False, variant=None):
        self._lib = lib(variant)
        c = _RefCfg()
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
        else:
            c.cfr_enable, c.cfr_clip, c.cfr_errclip = 0, 1.0, 1.0
        if tii:
            c.tii_enable = 1
            c.tii_comb, c.tii_pattern = tii[0], tii[1]
            c.tii_old_variant = int(tii[2]) if len(tii) > 2 else 0
        c.poly_threads = poly_threads
        enc = lambda s: s.encode() if s else None
        c.fir_taps_file = enc(fir_taps_file)
        c.poly_coef_file = enc(poly_coef_file)
        c.format = enc(fmt)
        c.stop_after = enc(stop_after)
        c.fixed_point = 1 if fixed_point else 0
        self._cfg = c
        self.mode = mode
        self._h = self._lib.ref_create(ctypes.byref(c))
        if not self._h:
            raise RuntimeError("ref_create: " + self._lib.ref_last_error().decode())
        self.latency = self._lib.ref_latency(self._h)
        self._out = np.empty(64 << 20, np.uint8)

    def feed(self, bits):
        """Feed one TF; returns raw output bytes (np.uint8, possibly empty)."""
        bits = np.ascontiguousarray(bits, np.uint8)
        n = self._lib.ref_process(self._h, bits.ctypes.data, bits.size,
                              self._out.ctypes.data, self._out.size)
        if n < 0:
            raise RuntimeError("ref_process: " + self._lib.ref_last_error().decode())
        return self._out[:n].copy()

    def feed_raw(self, bits):
        """Like feed() but leaves the result in the internal buffer (timing loops)."""
        n = self._lib.ref_process(self._h, bits.ctypes.data, bits.size,
                              self._out.ctypes.data, self._out.size)
        if n < 0:
            raise RuntimeError("ref_process: " + self._lib.ref_last_error().decode())
        return n

    def run(self, bits_tfs, dtype=np.complex64):
        """Feed a (n_tf, tf_bytes) array; flush the pipelined stages by feeding
        `latency` extra copies of the last TF; returns a list of n_tf arrays."""
        bits_tfs = np.ascontiguousarray(bits_tfs, np.uint8)
        outs = []
        for i in range(bits_tfs.shape[0] + self.latency):
            o = self.feed(bits_tfs[min(i, bits_tfs.shape[0] - 1)])
            if i >= self.latency:
                assert o.size, "reference produced no output"
                outs.append(o.view(dtype))
            else:
                assert o.size == 0
        return outs

    def get_param(self, name):
        """OfdmGeneratorCF32::get_parameter of this chain ("clip_stats", "papr", ...)."""
        buf = ctypes.create_string_buffer(512)
        if self._lib.ref_ofdm_get_parameter(self._h, name.encode(), buf, 512) != 0:
            raise RuntimeError(self._lib.ref_last_error().decode())
        return buf.value.decode()

    def close(self):
        if self._h:
            self._lib.ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# Channel coding: the reference's EtiReader + coding graph,
# oracle/ref_coder_harness.cpp
# ---------------------------------------------------------------------------
class RefCoder:
    """Raw ETI(NI) frames in, BlockPartitioner blocks out, through the unmodified reference."""

    def __init__(self):
        L = lib()
        L.refc_create.restype = ctypes.c_void_p
        L.refc_destroy.argtypes = [ctypes.c_void_p]
        L.refc_last_error.restype = ctypes.c_char_p
        L.refc_feed.restype = ctypes.c_long
        L.refc_feed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.refc_describe.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_int]
        self._h = L.refc_create()
        if not self._h:
            raise RuntimeError(L.refc_last_error().decode())
        self._out = np.zeros(4 * (384 + 6912), np.uint8)

    def feed(self, frame):
        frame = np.ascontiguousarray(frame, np.uint8)
        n = lib().refc_feed(self._h, frame.ctypes.data, frame.size, self._out.ctypes.data, self._out.size)
        if n < 0:
            raise RuntimeError(lib().refc_last_error().decode())
        return self._out[:n].copy() if n else None

    def run(self, frames):
        out = [self.feed(f) for f in np.ascontiguousarray(frames, np.uint8).reshape(-1, 6144)]
        return [o for o in out if o is not None]

    def describe(self):
        """[(framesize, out_bytes, start_cu, ((length, pattern), ...)), ...]: FIC first, as the
        reference's FicSource / SubchannelSource objects report it (after the first frame)."""
        cap_s, cap_r = 65, 512
        fs, ob, st, nr = (np.zeros(cap_s, np.uint32) for _ in range(4))
        rules = np.zeros(2 * cap_r, np.uint32)
        n = lib().refc_describe(self._h, fs.ctypes.data, ob.ctypes.data, st.ctypes.data, nr.ctypes.data,
                                rules.ctypes.data, cap_s, cap_r)
        if n < 0:
            raise RuntimeError(lib().refc_last_error().decode())
        out, r = [], 0
        for i in range(n):
            rl = tuple((int(rules[2 * (r + k)]), int(rules[2 * (r + k) + 1])) for k in range(nr[i]))
            r += int(nr[i])
            out.append((int(fs[i]), int(ob[i]), int(st[i]), rl))
        return out

    def close(self):
        if self._h:
            lib().refc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
