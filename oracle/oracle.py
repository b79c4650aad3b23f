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
    ]


def build():
    """Compile liboracle.so (gcc, ~1 s). Building the checker is not using it."""
    src = os.path.join(HERE, "dabmod_oracle.c")
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < os.path.getmtime(src):
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
                 cfr=None, tii=None, fir_taps=None, poly=None, lut=None, fmt=None):
        c = Cfg()
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
        dt = FORMAT_DTYPE[self.cfg.format] if st == 0 else np.complex64
        return self._buf[:n].copy().view(dt)

    def run(self, bits_tfs, stage="final"):
        return [self.process(b, stage) for b in np.ascontiguousarray(bits_tfs, np.uint8)]

    @property
    def clipped(self):
        return lib().dabo_chain_clipped(self._h)

    def close(self):
        if self._h:
            lib().dabo_chain_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
