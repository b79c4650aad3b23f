#!/usr/bin/env python
"""Golden blocks for the channel coding (row N1): BlockPartitioner output of the UNMODIFIED
reference (oracle/_ref, needs /root/reference) for the synthetic multiplexes of tests/test_coder.py.
Stored compactly: the last two transmission frames (full time-interleaver memory) and a byte sum per TF.

    python tests/golden/make_coder_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_coder  # noqa: E402
from oracle import refwrap  # noqa: E402

out = {}
for name, (mode, subch) in sorted(test_coder.multiplexes().items()):
    cif = {1: 4, 2: 1, 3: 1, 4: 2}[mode]
    n_frames, seed = 24 * cif, 1000 + len(name)
    frames = test_coder.eti_mod().synth_eti(mode, subch, n_frames, seed=seed)
    blocks = np.stack(refwrap.RefCoder().run(frames))
    out[name + "/n_frames"] = n_frames
    out[name + "/seed"] = seed
    out[name + "/shape"] = np.array(blocks.shape)
    out[name + "/last2"] = blocks[-2:]
    out[name + "/rowsum"] = blocks.astype(np.uint64).sum(axis=1)
np.savez_compressed(os.path.join(HERE, "coder_blocks.npz"), **out)
print("wrote coder_blocks.npz:", {k: v.shape for k, v in out.items() if k.endswith("last2")})
