"""ETI(NI) helpers on the host: a synthetic multiplex generator for tests and
bench.py, following SURVEY.md section 8(d) and the frame layout the reference's
EtiReader walks through (src/EtiReader.cpp:93-284, src/Eti.h:50-100).

No compute for the product happens here.
"""
import numpy as np

ETI_FRAME = 6144
MID_OF_MODE = {1: 1, 2: 2, 3: 3, 4: 0}


def eep_tpl(option, level):
    """TPL byte of an EEP subchannel: long form flag, option (0 = A, 1 = B), level 1..4."""
    return 0x20 | ((option & 7) << 2) | ((level - 1) & 3)


def synth_eti(mode, subchannels, n_frames, seed=1234):
    """n_frames raw ETI(NI) frames, (n_frames, 6144) uint8.

    subchannels: [(start_address_cu, stl (64-bit words per frame), tpl), ...].
    FIC and MST payload bytes are uniform random (CRCs are not checked by the
    reference's EtiReader, src/EtiReader.cpp:190-249); frame phase starts at 0."""
    rng = np.random.default_rng(seed)
    nst = len(subchannels)
    ficlen = 128 if mode == 3 else 96
    mst = sum(stl * 8 for _, stl, _ in subchannels)
    fl = nst + 1 + (ficlen + mst) // 4
    used = 4 + 4 + 4 * nst + 4 + ficlen + mst + 4 + 4
    if used > ETI_FRAME:
        raise ValueError("multiplex does not fit an ETI(NI) frame")
    frames = np.full((n_frames, ETI_FRAME), 0x55, np.uint8)
    for n in range(n_frames):
        f = frames[n]
        f[0:4] = (0xFF, 0x07, 0x3A, 0xB6) if n % 2 == 0 else (0xFF, 0xF8, 0xC5, 0x49)
        f[4] = n % 250
        f[5] = 0x80 | nst
        f[6] = ((n % 8) << 5) | (MID_OF_MODE[mode] << 3) | ((fl >> 8) & 7)
        f[7] = fl & 0xFF
        for i, (sad, stl, tpl) in enumerate(subchannels):
            c = f[8 + 4 * i: 12 + 4 * i]
            c[0] = ((i & 0x3F) << 2) | ((sad >> 8) & 3)
            c[1] = sad & 0xFF
            c[2] = ((tpl & 0x3F) << 2) | ((stl >> 8) & 3)
            c[3] = stl & 0xFF
        o = 8 + 4 * nst
        f[o:o + 4] = (0xFF, 0xFF, 0x00, 0x00)
        o += 4
        f[o:o + ficlen + mst] = rng.integers(0, 256, ficlen + mst, dtype=np.uint8)
        o += ficlen + mst
        f[o:o + 4] = 0
        f[o + 4:o + 8] = 0xFF
    return frames


def synth_eti_range(mode, subchannels, first_frame, n_frames, seed=1234, block=256, out=None):
    """Frames [first_frame, first_frame + n_frames) of a long synthetic stream whose payload is seeded per block of
    `block` frames, so that every rank of a sharded run can generate exactly its own range (bench.py: 65 536 frames
    over 8 GPUs) and all ranks agree on the stream.  Same frame layout as synth_eti; headers depend on the frame
    number only (FCT = n mod 250, FP = n mod 8)."""
    if out is None:
        out = np.empty((n_frames, ETI_FRAME), np.uint8)
    nst = len(subchannels)
    ficlen = 128 if mode == 3 else 96
    mst = sum(stl * 8 for _, stl, _ in subchannels)
    o = 8 + 4 * nst + 4
    pos = 0
    while pos < n_frames:
        n = first_frame + pos
        b, within = divmod(n, block)
        tmpl = synth_eti(mode, subchannels, block, seed=(seed * 1000003 + b) & 0x7fffffff)
        idx = np.arange(b * block, (b + 1) * block)
        tmpl[:, 0:4] = np.where((idx % 2 == 0)[:, None], np.array([0xFF, 0x07, 0x3A, 0xB6], np.uint8),
                                np.array([0xFF, 0xF8, 0xC5, 0x49], np.uint8))
        tmpl[:, 4] = idx % 250
        tmpl[:, 6] = (tmpl[:, 6] & 0x1F) | ((idx % 8) << 5).astype(np.uint8)
        take = min(block - within, n_frames - pos)
        out[pos:pos + take] = tmpl[within:within + take]
        pos += take
    assert o + ficlen + mst + 8 <= ETI_FRAME
    return out


def default_multiplex():
    """six 128 kbit/s EEP 3-A subchannels (96 CU each)."""
    return [(96 * i, 48, eep_tpl(0, 3)) for i in range(6)]
