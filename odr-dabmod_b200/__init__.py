"""Python binding of libdabmod_b200.so (include/dabmod_b200.h) via ctypes.

This is harness code for tests and bench.py: the product is the CUDA shared
library.  There is deliberately NO fallback: if the library is missing or no
B200 is visible, construction raises.

Directory name contains a hyphen (it is the repo's package directory, not an
importable name); load it with `dabmod_loader.load()` from the repo root.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# DABMOD_B200_LIB: an experimental build of the same library (tools/*_exp.sh)
LIB_PATH = os.environ.get("DABMOD_B200_LIB") or os.path.join(HERE, "libdabmod_b200.so")
CSRC = os.path.join(HERE, "csrc")

ABI_VERSION = 2
GAIN_MODES = {"fix": 0, "max": 1, "var": 2}
FORMATS = {None: 0, "": 0, "complexf": 0, "s16": 1, "u8": 2, "s8": 3}
FORMAT_DTYPE = {0: np.complex64, 1: np.int16, 2: np.uint8, 3: np.int8}
FORMAT_BYTES = {0: 8, 1: 4, 2: 2, 3: 2}

# every symbol include/dabmod_b200.h declares
EXPORTS = [
    "dabmod_b200_config_init", "dabmod_b200_default_fir_taps", "dabmod_b200_create",
    "dabmod_b200_destroy", "dabmod_b200_tf_in_bytes", "dabmod_b200_tf_out_bytes",
    "dabmod_b200_tf_out_samples", "dabmod_b200_process", "dabmod_b200_process_batch",
    "dabmod_b200_process_batch_device", "dabmod_b200_process_batch_to_fd", "dabmod_b200_synchronize", "dabmod_b200_reset",
    "dabmod_b200_seek", "dabmod_b200_set_param", "dabmod_b200_get_param",
    "dabmod_b200_num_clipped_samples", "dabmod_b200_last_launch_count", "dabmod_b200_last_error",
    "dabmod_b200_table_interleaver", "dabmod_b200_table_phase_ref", "dabmod_b200_table_tii",
    "dabmod_b200_table_cic", "dabmod_b200_resampler_sizes", "dabmod_b200_kernel_time",
    "dabmod_b200_device_out", "dabmod_b200_host_register", "dabmod_b200_host_unregister",
    # row N1: channel coding
    "dabmod_b200_eti_describe", "dabmod_b200_coder_create", "dabmod_b200_coder_destroy",
    "dabmod_b200_coder_tf_bytes", "dabmod_b200_coder_frames_per_tf", "dabmod_b200_coder_stream_offset",
    "dabmod_b200_coder_process",
    "dabmod_b200_coder_process_device", "dabmod_b200_coder_reset", "dabmod_b200_coder_prime",
    "dabmod_b200_process_eti_batch", "dabmod_b200_process_eti_batch_to_fd", "dabmod_b200_seek_eti",
    "dabmod_b200_coder_last_error",
]


class Config(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_uint32),
        ("device", ctypes.c_int32),
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
        ("dpd_mode", ctypes.c_int32),
        ("dpd_coefs", ctypes.POINTER(ctypes.c_float)),
        ("format", ctypes.c_int32),
        ("max_batch", ctypes.c_int32),
        ("fft_engine", ctypes.c_int32),
    ]


class Rule(ctypes.Structure):
    _fields_ = [("length", ctypes.c_uint32), ("pattern", ctypes.c_uint32)]


class Stream(ctypes.Structure):
    """dabmod_b200_stream"""
    _fields_ = [("framesize", ctypes.c_uint32), ("out_bytes", ctypes.c_uint32), ("start_cu", ctypes.c_uint32),
                ("n_rules", ctypes.c_uint32), ("rules", Rule * 8)]

    def as_tuple(self):
        return (self.framesize, self.out_bytes, self.start_cu,
                tuple((self.rules[i].length, self.rules[i].pattern) for i in range(self.n_rules)))


class DabModError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dabmod_b200 error %d: %s" % (code, msg))
        self.code = code


def build(verbose=False):
    """Compile libdabmod_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC] + ([] if verbose else ["-s"]))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    L.dabmod_b200_config_init.argtypes = [ctypes.POINTER(Config)]
    L.dabmod_b200_default_fir_taps.argtypes = [vp, ctypes.c_int]
    L.dabmod_b200_create.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(vp)]
    L.dabmod_b200_destroy.argtypes = [vp]
    for f in ("tf_in_bytes", "tf_out_bytes", "tf_out_samples"):
        getattr(L, "dabmod_b200_" + f).restype = sz
        getattr(L, "dabmod_b200_" + f).argtypes = [vp]
    L.dabmod_b200_process.argtypes = [vp, vp, sz, vp, sz, ctypes.POINTER(sz)]
    L.dabmod_b200_process_batch.argtypes = [vp, vp, sz, vp, sz, ctypes.POINTER(sz)]
    L.dabmod_b200_process_batch_device.argtypes = [vp, vp, sz, vp, vp]
    L.dabmod_b200_process_batch_to_fd.argtypes = [vp, vp, sz, ctypes.c_int, ctypes.POINTER(sz)]
    L.dabmod_b200_synchronize.argtypes = [vp]
    L.dabmod_b200_reset.argtypes = [vp]
    L.dabmod_b200_seek.argtypes = [vp, ctypes.c_uint64, vp, sz]
    L.dabmod_b200_set_param.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p]
    L.dabmod_b200_get_param.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, sz]
    L.dabmod_b200_num_clipped_samples.restype = ctypes.c_uint64
    L.dabmod_b200_num_clipped_samples.argtypes = [vp]
    L.dabmod_b200_last_launch_count.restype = ctypes.c_uint32
    L.dabmod_b200_last_launch_count.argtypes = [vp]
    L.dabmod_b200_last_error.restype = ctypes.c_char_p
    L.dabmod_b200_table_interleaver.argtypes = [ctypes.c_int, vp, ctypes.c_int]
    L.dabmod_b200_table_phase_ref.argtypes = [ctypes.c_int, vp, ctypes.c_int]
    L.dabmod_b200_table_tii.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int]
    L.dabmod_b200_table_cic.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, vp]
    L.dabmod_b200_resampler_sizes.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.dabmod_b200_kernel_time.argtypes = [vp, ctypes.c_int, ctypes.c_char_p, sz, ctypes.POINTER(ctypes.c_float)]
    L.dabmod_b200_device_out.restype = vp
    L.dabmod_b200_device_out.argtypes = [vp]
    L.dabmod_b200_eti_describe.argtypes = [vp, sz, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(Stream), ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_int)]
    L.dabmod_b200_coder_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(Stream), ctypes.c_int,
                                           ctypes.c_int, ctypes.POINTER(vp)]
    L.dabmod_b200_coder_destroy.argtypes = [vp]
    L.dabmod_b200_coder_tf_bytes.restype = sz
    L.dabmod_b200_coder_tf_bytes.argtypes = [vp]
    L.dabmod_b200_coder_frames_per_tf.argtypes = [vp]
    L.dabmod_b200_coder_process.argtypes = [vp, vp, sz, vp, sz, ctypes.POINTER(sz)]
    L.dabmod_b200_coder_process_device.argtypes = [vp, vp, sz, vp, vp]
    L.dabmod_b200_coder_reset.argtypes = [vp]
    L.dabmod_b200_coder_prime.argtypes = [vp, vp, sz]
    L.dabmod_b200_process_eti_batch.argtypes = [vp, vp, vp, sz, vp, sz, ctypes.POINTER(sz)]
    L.dabmod_b200_process_eti_batch_to_fd.argtypes = [vp, vp, vp, sz, ctypes.c_int, ctypes.POINTER(sz)]
    L.dabmod_b200_seek_eti.argtypes = [vp, vp, ctypes.c_uint64, vp, sz]
    L.dabmod_b200_coder_last_error.restype = ctypes.c_char_p
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise DabModError(rc, lib().dabmod_b200_last_error().decode(errors="replace"))


def resampler_sizes(in_rate, out_rate, resolution):
    """(Ni, No) the reference's Resampler picks for this ratio (Resampler.cpp:65-76)."""
    ni, no = ctypes.c_int(), ctypes.c_int()
    _check(lib().dabmod_b200_resampler_sizes(in_rate, out_rate, resolution, ctypes.byref(ni), ctypes.byref(no)))
    return ni.value, no.value


def default_fir_taps():
    n = lib().dabmod_b200_default_fir_taps(None, 0)
    t = np.zeros(n, np.float32)
    lib().dabmod_b200_default_fir_taps(t.ctypes.data, n)
    return t


def load_fir_taps(taps_file):
    """FIRFilter::load_filter_taps (reference FIRFilter.cpp:95-141): "default" or a
    text file `ntaps` followed by the taps."""
    if taps_file == "default":
        return default_fir_taps()
    with open(taps_file) as f:
        tok = f.read().split()
    if not tok:
        raise RuntimeError("FIRFilter: taps file has invalid format.")
    n = int(tok[0])
    if n <= 0:
        raise RuntimeError("FIRFilter: taps file has invalid format.")
    if len(tok) < 1 + n:
        raise RuntimeError("FIRFilter: file %s should contain %d taps, but EOF reached after %d taps!"
                           % (taps_file, n, len(tok) - 1))
    return np.array([float(x) for x in tok[1:1 + n]], np.float32)


def load_dpd_coefs(coef_file):
    """MemlessPoly::load_coefficients (reference MemlessPoly.cpp:145-235).
    Returns (dpd_mode, float32 array) in the layout dabmod_b200_config.dpd_coefs wants."""
    with open(coef_file) as f:
        tok = f.read().split()
    fmt = int(tok[0])
    if fmt == 1:
        n = int(tok[1])
        if n != 5:
            raise RuntimeError("MemlessPoly: invalid number of coefs: %d expected 5" % n)
        vals = [float(x) for x in tok[2:12]]
        if len(vals) != 10:
            raise RuntimeError("MemlessPoly: coefs file invalid !")
        return 1, np.array(vals, np.float32)
    if fmt == 2:
        vals = [float(x) for x in tok[1:34]]
        if len(vals) != 33:
            raise RuntimeError("MemlessPoly: coefs file invalid !")
        return 2, np.array(vals, np.float32)
    raise RuntimeError("MemlessPoly: coef file has unknown format %d" % fmt)


class Modulator:
    """One modulator stream on one GPU (a dabmod_b200 handle)."""

    def __init__(self, mode=1, gain_mode="var", output_rate=2048000, clock_rate=0,
                 digital_gain=1.0, normalise=1.0, gain_variance=4.0, window_overlap=0,
                 cfr=None, tii=None, fir_taps=None, poly=None, lut=None, fmt=None,
                 max_batch=1, device=0, fixed_point=False):
        L = lib()
        c = Config()
        L.dabmod_b200_config_init(ctypes.byref(c))
        c.fft_engine = 1 if fixed_point else 0
        c.device = device
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
        keep = []
        if fir_taps is not None:
            if isinstance(fir_taps, str):
                fir_taps = load_fir_taps(fir_taps)
            t = np.ascontiguousarray(fir_taps, np.float32)
            keep.append(t)
            c.fir_ntaps = t.size
            c.fir_taps = t.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        if poly is not None:
            p = np.ascontiguousarray(poly, np.float32)
            assert p.size == 10
            keep.append(p)
            c.dpd_mode, c.dpd_coefs = 1, p.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        elif lut is not None:
            p = np.ascontiguousarray(np.r_[np.float32(lut[0]), np.asarray(lut[1], np.float32)], np.float32)
            assert p.size == 33
            keep.append(p)
            c.dpd_mode, c.dpd_coefs = 2, p.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        c.format = FORMATS[fmt]
        c.max_batch = max_batch
        self.cfg = c
        self._h = ctypes.c_void_p()
        _check(L.dabmod_b200_create(ctypes.byref(c), ctypes.byref(self._h)))
        self.tf_in_bytes = L.dabmod_b200_tf_in_bytes(self._h)
        self.tf_out_bytes = L.dabmod_b200_tf_out_bytes(self._h)
        self.tf_out_samples = L.dabmod_b200_tf_out_samples(self._h)
        self.out_dtype = np.int16 if fixed_point else FORMAT_DTYPE[c.format]
        self.max_batch = max_batch

    # -- host buffers -------------------------------------------------------
    def process(self, bits):
        """One TF: ModCodec::process(Buffer*, Buffer*) semantics. Returns a numpy array."""
        bits = np.ascontiguousarray(bits, np.uint8)
        out = np.empty(self.tf_out_bytes, np.uint8)
        n = ctypes.c_size_t()
        _check(lib().dabmod_b200_process(self._h, bits.ctypes.data, bits.size, out.ctypes.data,
                                         out.size, ctypes.byref(n)))
        return out[:n.value].view(self.out_dtype)

    def process_batch(self, bits, out=None):
        """bits: (n_tf, tf_in_bytes) uint8. Returns (n_tf, samples[*2]) array."""
        bits = np.ascontiguousarray(bits, np.uint8)
        n_tf = bits.size // self.tf_in_bytes
        assert bits.size == n_tf * self.tf_in_bytes
        if out is None:
            out = np.empty(n_tf * self.tf_out_bytes, np.uint8)
        n = ctypes.c_size_t()
        _check(lib().dabmod_b200_process_batch(self._h, bits.ctypes.data, n_tf, out.ctypes.data,
                                               out.nbytes, ctypes.byref(n)))
        flat = out.reshape(-1)[:n.value].view(self.out_dtype)
        return flat.reshape(n_tf, flat.size // n_tf if n_tf else 0)

    def process_batch_ptr(self, bits_ptr, n_tf, out_ptr, out_cap):
        """Raw host pointers (e.g. pinned torch tensors)."""
        n = ctypes.c_size_t()
        _check(lib().dabmod_b200_process_batch(self._h, bits_ptr, n_tf, out_ptr, out_cap, ctypes.byref(n)))
        return n.value

    # -- device buffers -----------------------------------------------------
    def process_batch_to_fd(self, bits, fd):
        """n_tf TFs of the stream written to the file descriptor `fd` (the reference's OutputFile sink) through the
        handle's pinned ring; returns the number of bytes written."""
        bits = np.ascontiguousarray(bits, np.uint8).reshape(-1, self.tf_in_bytes)
        n = ctypes.c_size_t(0)
        _check(lib().dabmod_b200_process_batch_to_fd(self._h, bits.ctypes.data, bits.shape[0], fd, ctypes.byref(n)))
        return n.value

    def process_batch_device(self, d_bits_ptr, n_tf, d_out_ptr, stream=0):
        _check(lib().dabmod_b200_process_batch_device(self._h, d_bits_ptr, n_tf, d_out_ptr, stream or None))

    def synchronize(self):
        _check(lib().dabmod_b200_synchronize(self._h))

    def reset(self):
        _check(lib().dabmod_b200_reset(self._h))

    def seek(self, tf_index, prev_bits=None):
        if prev_bits is None:
            _check(lib().dabmod_b200_seek(self._h, tf_index, None, 0))
        else:
            pb = np.ascontiguousarray(prev_bits, np.uint8)
            _check(lib().dabmod_b200_seek(self._h, tf_index, pb.ctypes.data, pb.size))

    def set_param(self, name, value):
        _check(lib().dabmod_b200_set_param(self._h, name.encode(), str(value).encode()))

    def get_param(self, name):
        buf = ctypes.create_string_buffer(256)
        _check(lib().dabmod_b200_get_param(self._h, name.encode(), buf, 256))
        return buf.value.decode()

    def kernel_times(self):
        """[(kernel name, milliseconds)] of the last process call; needs set_param("profile", 1)."""
        out = []
        for i in range(self.last_launch_count):
            name = ctypes.create_string_buffer(64)
            ms = ctypes.c_float()
            _check(lib().dabmod_b200_kernel_time(self._h, i, name, 64, ctypes.byref(ms)))
            out.append((name.value.decode(), ms.value))
        return out

    @property
    def num_clipped_samples(self):
        return lib().dabmod_b200_num_clipped_samples(self._h)

    @property
    def last_launch_count(self):
        return lib().dabmod_b200_last_launch_count(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().dabmod_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# Row N1: channel coding (ETI frames -> transmission-frame blocks)
# ---------------------------------------------------------------------------
ETI_FRAME = 6144


def _check_coder(rc):
    if rc != 0:
        raise DabModError(rc, lib().dabmod_b200_coder_last_error().decode(errors="replace"))


def make_streams(desc):
    """[(framesize, out_bytes, start_cu, ((length, pattern), ...)), ...] -> dabmod_b200_stream array"""
    arr = (Stream * len(desc))()
    for i, d in enumerate(desc):
        fs, ob, sc, rules = d.as_tuple() if isinstance(d, Stream) else d
        arr[i].framesize, arr[i].out_bytes, arr[i].start_cu, arr[i].n_rules = fs, ob, sc, len(rules)
        for k, (a, b) in enumerate(rules):
            arr[i].rules[k] = Rule(a, b)
    return arr


def eti_describe(frame):
    """(mode, [stream tuples]) from one ETI(NI) frame header: dabmod_b200_eti_describe (host only)."""
    frame = np.ascontiguousarray(frame, np.uint8)
    st = (Stream * 65)()
    mode, n = ctypes.c_int(), ctypes.c_int()
    _check_coder(lib().dabmod_b200_eti_describe(frame.ctypes.data, frame.size, ctypes.byref(mode), st, 65,
                                                ctypes.byref(n)))
    return mode.value, [st[i].as_tuple() for i in range(n.value)]


class Coder:
    """One multiplex configuration on one GPU (a dabmod_b200_coder handle)."""

    def __init__(self, mode, streams, max_frames=4, device=0):
        arr = make_streams(streams)
        self._h = ctypes.c_void_p()
        _check_coder(lib().dabmod_b200_coder_create(device, mode, arr, len(arr), max_frames, ctypes.byref(self._h)))
        self.tf_bytes = lib().dabmod_b200_coder_tf_bytes(self._h)
        self.frames_per_tf = lib().dabmod_b200_coder_frames_per_tf(self._h)
        self.max_frames = max_frames

    def process(self, frames):
        """frames: (n, 6144) uint8, n a multiple of frames_per_tf -> (n / frames_per_tf, tf_bytes)"""
        frames = np.ascontiguousarray(frames, np.uint8)
        n = frames.size // ETI_FRAME
        out = np.empty((n // max(self.frames_per_tf, 1)) * self.tf_bytes, np.uint8)
        nb = ctypes.c_size_t()
        _check_coder(lib().dabmod_b200_coder_process(self._h, frames.ctypes.data, n, out.ctypes.data, out.size,
                                                     ctypes.byref(nb)))
        return out[:nb.value].reshape(-1, self.tf_bytes)

    def process_device(self, d_eti_ptr, n_frames, d_bits_ptr, stream=0):
        _check_coder(lib().dabmod_b200_coder_process_device(self._h, d_eti_ptr, n_frames, d_bits_ptr, stream or None))

    def prime(self, frames):
        frames = np.ascontiguousarray(frames, np.uint8)
        _check_coder(lib().dabmod_b200_coder_prime(self._h, frames.ctypes.data, frames.size // ETI_FRAME))

    def reset(self):
        _check_coder(lib().dabmod_b200_coder_reset(self._h))

    def modulate(self, modulator, frames):
        """ETI frames -> I/Q through coder + modulator chained on the device."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n = frames.size // ETI_FRAME
        n_tf = n // self.frames_per_tf
        out = np.empty(n_tf * modulator.tf_out_bytes, np.uint8)
        nb = ctypes.c_size_t()
        _check_coder(lib().dabmod_b200_process_eti_batch(modulator._h, self._h, frames.ctypes.data, n,
                                                         out.ctypes.data, out.size, ctypes.byref(nb)))
        flat = out[:nb.value].view(modulator.out_dtype)
        return flat.reshape(n_tf, flat.size // n_tf if n_tf else 0)

    def modulate_ptr(self, modulator, eti_ptr, n_frames, out_ptr, out_cap):
        """Raw host pointers (e.g. pinned torch tensors): dabmod_b200_process_eti_batch."""
        nb = ctypes.c_size_t()
        _check_coder(lib().dabmod_b200_process_eti_batch(modulator._h, self._h, eti_ptr, n_frames, out_ptr, out_cap,
                                                         ctypes.byref(nb)))
        return nb.value

    def modulate_to_fd(self, modulator, frames, fd):
        """ETI frames -> I/Q written to the descriptor `fd` (the reference's OutputFile sink); returns bytes written."""
        frames = np.ascontiguousarray(frames, np.uint8)
        nb = ctypes.c_size_t()
        _check_coder(lib().dabmod_b200_process_eti_batch_to_fd(modulator._h, self._h, frames.ctypes.data,
                                                               frames.size // ETI_FRAME, fd, ctypes.byref(nb)))
        return nb.value

    def seek(self, modulator, tf_index, frames_before):
        """Position coder + modulator at transmission frame `tf_index` of the stream; `frames_before` = the ETI frames
        that precede it (at least min(tf_index * frames_per_tf, 15 + frames_per_tf) of them)."""
        fb = np.ascontiguousarray(frames_before, np.uint8)
        _check_coder(lib().dabmod_b200_seek_eti(modulator._h, self._h, tf_index, fb.ctypes.data if fb.size else None,
                                                fb.size // ETI_FRAME))

    def close(self):
        if getattr(self, "_h", None):
            lib().dabmod_b200_coder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
