import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libdabmod_ref.so (the compiled reference)")


def rel_rms(a, b):
    """||a - b||_2 / ||b||_2 in float64 (SURVEY.md section 8(d) parity metric)."""
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    den = np.sqrt(np.sum(np.abs(b) ** 2))
    num = np.sqrt(np.sum(np.abs(a - b) ** 2))
    return num / den if den > 0 else num


@pytest.fixture(scope="session")
def rng():
    return np.random.default_rng(20261017)


def write_taps_file(path, taps):
    """FIRFilter taps file format (FIRFilter.cpp:103-133): count, then one tap per line."""
    with open(path, "w") as f:
        f.write("%d\n" % len(taps))
        for t in taps:
            f.write("%.9g\n" % float(t))


def write_poly_file(path, am, pm):
    """MemlessPoly odd-polynomial file (MemlessPoly.cpp:145-202): 1, 5, am0..4, pm0..4."""
    with open(path, "w") as f:
        f.write("1\n5\n")
        for c in list(am) + list(pm):
            f.write("%.9g\n" % float(c))


def write_lut_file(path, scale, lut):
    """MemlessPoly LUT file as the loader reads it (MemlessPoly.cpp:203-229): 2, scalefactor, 32 entries."""
    with open(path, "w") as f:
        f.write("2\n%.9g\n" % float(scale))
        for c in lut:
            f.write("%.9g\n" % float(c))
