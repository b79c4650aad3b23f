#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference code.

Run in the build container (needs /root/reference):
    python -c "import __graft_entry__ as g; g.build()"     # builds oracle/_ref/libdabmod_ref.so
    python tests/golden/make_golden.py

Each fixture holds the input blocks (`bits`, what BlockPartitioner hands to
QpskSymbolMapper) and, per TF, slices of the reference's final output buffer
(what OutputMemory copies out): head, tail, every STRIDE-th sample, plus
float64 checksums over the whole TF.  Slices keep the fixtures small; the
checksums cover the samples the slices skip.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refwrap                     # noqa: E402
from conftest import write_poly_file           # noqa: E402
from golden_cases import CASES, HEAD, STRIDE, slices   # noqa: E402


def main():
    tmp = tempfile.mkdtemp()
    only = sys.argv[1] if len(sys.argv) > 1 else ""      # optional name prefix: regenerate a subset
    for name, case in CASES.items():
        if not name.startswith(only):
            continue
        kw = dict(case["cfg"])
        rng = np.random.default_rng(case["seed"])
        bits = rng.integers(0, 256, (case["n_tf"], refwrap.TF_BYTES[kw["mode"]]), dtype=np.uint8)
        ref_kw = dict(kw)
        if ref_kw.pop("fir", False):
            ref_kw["fir_taps_file"] = "default"
        poly = ref_kw.pop("poly", None)
        if poly is not None:
            p = os.path.join(tmp, name + ".coef")
            write_poly_file(p, poly[:5], poly[5:])
            ref_kw["poly_coef_file"] = p
            ref_kw["poly_threads"] = 1
        dt = {None: np.complex64, "s16": np.int16, "u8": np.uint8, "s8": np.int8}[kw.get("fmt")]
        if kw.get("fixed_point"):
            dt = np.int16
        outs = refwrap.RefChain(**ref_kw).run(bits, dtype=dt)
        arrs = {"bits": bits}
        for i, o in enumerate(outs):
            h, t, s, chk = slices(o)
            arrs["head%d" % i], arrs["tail%d" % i], arrs["stride%d" % i], arrs["chk%d" % i] = h, t, s, chk
            arrs["size%d" % i] = np.array([o.size])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrs)
        print("%-28s %d TFs  %7.1f KB" % (name, len(outs), os.path.getsize(path) / 1024))
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "ODR-DabMod v3.0.1 "
                   "(87fce948967230707de283b3e7d0e833f00bfad2), FFTW replaced by the vendored KISS FFT as float "
                   "(oracle/refshim)", "head": HEAD, "stride": STRIDE, "cases": sorted(CASES)}, f, indent=1)


if __name__ == "__main__":
    main()
