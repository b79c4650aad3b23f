"""The register butterflies of the CUDA kernels (fft.cuh, generated fft_reg.cuh),
compiled for the host by nvcc and checked against numpy on the CPU."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "fft_host.cu")
LIB = os.path.join(HERE, "host", "libfft_host.so")


@pytest.fixture(scope="module")
def lib():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    deps = [SRC] + [os.path.join(HERE, "..", "odr-dabmod_b200", "csrc", f) for f in ("fft.cuh", "fft_reg.cuh", "resample_q.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-o", LIB, SRC])
    L = ctypes.CDLL(LIB)
    L.fft_host.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.rq_hop_host.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return L


def test_generated_header_is_current():
    gen = subprocess.run(["python", os.path.join(HERE, "..", "tools", "gen_fft_reg.py")], capture_output=True,
                         text=True, check=True).stdout
    with open(os.path.join(HERE, "..", "odr-dabmod_b200", "csrc", "fft_reg.cuh")) as f:
        assert f.read() == gen


@pytest.mark.parametrize("inv", [0, 1])
@pytest.mark.parametrize("n", [4, 5, 8, -8, -10, 16, 20, 32, 64])
def test_register_butterflies(lib, n, inv):
    rng = np.random.default_rng(100 + abs(n))
    size = abs(n)
    x = (rng.standard_normal(size) + 1j * rng.standard_normal(size)).astype(np.complex64)
    io = x.copy()
    assert lib.fft_host(n, inv, io.ctypes.data) == 0
    want = np.fft.ifft(x.astype(np.complex128)) * size if inv else np.fft.fft(x.astype(np.complex128))
    keep = 5 if n == -10 else size      # fft10_lo returns the first five bins
    err = np.linalg.norm(io[:keep] - want[:keep]) / np.linalg.norm(want[:keep])
    assert err < 3e-7, err
    # unit impulses: every output bin of every input position
    for pos in (1, size - 1):
        e = np.zeros(size, np.complex64)
        e[pos] = 1
        io = e.copy()
        lib.fft_host(n, inv, io.ctypes.data)
        k = np.arange(size)
        want = np.exp((1 if inv else -1) * 2j * np.pi * k * pos / size)
        assert np.abs(io - want)[:keep].max() < 5e-7


@pytest.mark.parametrize("P", [2, 3, 4, 5])
def test_resample_q_hop(lib, P):
    """One hop of k_resample_q (spectrum folding onto 4000 slots, radix 20 x 20 x 10 phase transforms, phase
    interleaving) run thread by thread on the host against the plain No-point inverse transform of the re-laid-out
    spectrum (Resampler.cpp:153-183)."""
    ni, no = 4096, 4000 * P
    rng = np.random.default_rng(40 + P)
    F = (rng.standard_normal(ni) + 1j * rng.standard_normal(ni)).astype(np.complex64)
    tw = np.exp(2j * np.pi * np.arange(no) / no).astype(np.complex64)
    out = np.zeros(P * 2000, np.complex64)
    assert lib.rq_hop_host(P, F.ctypes.data, tw.ctypes.data, out.ctypes.data) == 0
    B = np.zeros(no, np.complex128)
    B[: ni // 2] = F[: ni // 2]
    B[no - ni // 2:] = F[ni // 2:]
    B[ni // 2] = F[ni // 2]
    want = (np.fft.ifft(B) * no)[: no // 2]
    assert np.isfinite(out).all()
    err = np.linalg.norm(out - want) / np.linalg.norm(want)
    assert err < 5e-7, err
