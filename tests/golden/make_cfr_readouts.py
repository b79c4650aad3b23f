#!/usr/bin/env python
"""Generates tests/golden/cfr_readouts.json from the UNMODIFIED reference: the strings
OfdmGeneratorCF32::get_parameter("clip_stats" / "papr") returns (OfdmGenerator.cpp:419-453)
after given numbers of transmission frames.  Needs /root/reference (oracle/_ref built).

    python tests/golden/make_cfr_readouts.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refwrap                                 # noqa: E402
from golden_cases import READOUT_CASES                     # noqa: E402


def main():
    out = {}
    for name, case in READOUT_CASES.items():
        rng = np.random.default_rng(case["seed"])
        bits = rng.integers(0, 256, (case["n_tf"], refwrap.TF_BYTES[case["cfg"]["mode"]]), dtype=np.uint8)
        ref = refwrap.RefChain(**case["cfg"])
        at = {}
        for i in range(case["n_tf"]):
            ref.feed(bits[i])
            if i + 1 in case["after"]:
                at[str(i + 1)] = {"clip_stats": ref.get_param("clip_stats"), "papr": ref.get_param("papr")}
        out[name] = at
        print(name, at[str(case["after"][-1])])
    with open(os.path.join(HERE, "cfr_readouts.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
