"""CPU checks of the drop-in boundary: libdabmod_b200.so loads, exports every function include/dabmod_b200.h declares,
the Python binding's EXPORTS list is that same set, and the entry points that need no GPU behave (tables, defaults,
the resampler geometry of Resampler.cpp:65-76); without a CUDA device `create` fails loudly -- there is no CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

import dabmod_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dabmod_b200.h")


def declared_functions():
    with open(HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(dabmod_b200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def dm():
    return dabmod_loader.load()


def test_library_exports_every_declared_symbol(dm):
    names = declared_functions()
    assert len(names) >= 30
    lib = dm.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(dm.EXPORTS) == names          # the binding's list is the header's list


def test_config_defaults_match_the_reference(dm):
    """ConfigParser.h:45-96"""
    c = dm.Config()
    dm.lib().dabmod_b200_config_init(ctypes.byref(c))
    assert c.abi_version == dm.ABI_VERSION == 2
    assert (c.mode, c.gain_mode, c.output_rate, c.clock_rate) == (1, 2, 2048000, 0)
    assert (c.digital_gain, c.normalise, c.gain_variance) == (1.0, 1.0, 4.0)
    assert (c.window_overlap, c.cfr_enable, c.tii_enable, c.fir_ntaps, c.dpd_mode, c.format, c.fft_engine) == (0,) * 7


def test_tables_without_a_gpu(dm):
    taps = dm.default_fir_taps()
    assert len(taps) == 45 and np.allclose(taps, taps[::-1])           # FIRFilter.cpp:59-71, symmetric
    assert dm.resampler_sizes(2048000, 8192000, 2048) == (4096, 16384)  # Resampler.cpp:65-76
    assert dm.resampler_sizes(2048000, 10000000, 2048) == (4096, 20000)
    assert dm.resampler_sizes(2048000, 1536000, 512) == (1024, 768)
    idx = (ctypes.c_int32 * 1536)()
    assert dm.lib().dabmod_b200_table_interleaver(1, idx, 1536) == 0
    assert sorted(idx) == list(range(1536))                             # a permutation of the carriers


def test_create_fails_loudly_without_a_gpu(dm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(dm.DabModError) as e:
        dm.Modulator(mode=1)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
