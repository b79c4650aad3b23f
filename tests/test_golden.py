"""Golden vectors generated from the UNMODIFIED reference (tests/golden/make_golden.py).

  * not gpu: the oracle (oracle/dabmod_oracle.c) against the fixtures -- this is
    what pins the oracle where /root/reference does not exist (the GPU box);
  * gpu: the CUDA path through the C ABI against the same fixtures.

Tolerance: 1e-5 relative RMS is the north-star bar; we assert 2e-6.  Integer
formats: off by at most one count on < 1 % of the components.
"""
import os

import numpy as np
import pytest

import dabmod_loader
from conftest import rel_rms
from golden_cases import CASES, slices
from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-6


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def chain_kwargs(cfg):
    kw = dict(cfg)
    if kw.pop("fir", False):
        kw["fir_taps"] = oracle.fir_default_taps()
    return kw


def check(name, outs):
    g = load(name)
    case = CASES[name]
    integer = case["cfg"].get("fmt") is not None
    for i, o in enumerate(outs):
        assert o.size == int(g["size%d" % i][0])
        h, t, s, chk = slices(o)
        for part, got in (("head", h), ("tail", t), ("stride", s)):
            want = g["%s%d" % (part, i)]
            if integer:
                d = np.abs(got.astype(np.int32) - want.astype(np.int32))
                assert d.max() <= 1 and np.count_nonzero(d) < 0.01 * d.size + 2
            elif np.abs(want).max() == 0:
                assert np.abs(got).max() == 0            # plain null symbol
            else:
                assert rel_rms(got, want) < TOL, (name, i, part)
        # whole-TF checksums: energy to 1e-5, sums to 1e-5 of sqrt(N * energy)
        want = g["chk%d" % i]
        scale = np.sqrt(o.size * want[2]) + 1.0
        if integer:
            assert abs(chk[2] - want[2]) <= 2e-4 * want[2]
        else:
            assert abs(chk[2] - want[2]) <= 1e-5 * want[2]
            assert abs(chk[0] - want[0]) <= 1e-5 * scale and abs(chk[1] - want[1]) <= 1e-5 * scale


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    g = load(name)
    outs = oracle.OracleChain(**chain_kwargs(CASES[name]["cfg"])).run(g["bits"])
    check(name, outs)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(name):
    dm = dabmod_loader.load()
    g = load(name)
    bits = g["bits"]
    try:
        mod = dm.Modulator(max_batch=bits.shape[0], **chain_kwargs(CASES[name]["cfg"]))
    except dm.DabModError as e:
        if e.code == -4:
            pytest.skip("not implemented on the GPU yet: %s" % e)
        raise
    check(name, list(mod.process_batch(bits)))
