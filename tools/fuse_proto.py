#!/usr/bin/env python
"""Timing of the fused symbol + FIR prototype against the two-kernel headline step (C2, 1024 TFs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dabmod_loader  # noqa: E402

dm = dabmod_loader.load()
n = 1024
for fuse in (0, 1, 0, 1):
    mod = dm.Modulator(mode=1, fir_taps="default", max_batch=n)
    mod.set_param("fuse_proto", fuse)
    mod.set_param("profile", 1)
    bits = torch.randint(0, 256, (n, mod.tf_in_bytes), dtype=torch.uint8).cuda()
    out = torch.empty(n * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    ts = []
    for _ in range(8):
        mod.process_batch_device(bits.data_ptr(), n, out.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        ts.append(mod.kernel_times())
    med = {}
    for k, _ in ts[0]:
        med[k] = float(np.median([dict(t)[k] for t in ts[3:]]))
    print("fuse=%d" % fuse, med, "total %.4f ms" % sum(med.values()), flush=True)
    mod.close()
