#!/usr/bin/env python
"""TM I + FIR default taps + TII, 1024 TFs: the step with fir_kernel = 2 (symbol kernel, TII fill, k_fir) and 3 (fused)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dabmod_loader  # noqa: E402

dm = dabmod_loader.load()
n = 1024
for fk in (2, 3, 2, 3):
    mod = dm.Modulator(mode=1, fir_taps="default", tii=(3, 20), max_batch=n)
    mod.set_param("fir_kernel", fk)
    mod.set_param("profile", 1)
    bits = torch.randint(0, 256, (n, mod.tf_in_bytes), dtype=torch.uint8).cuda()
    out = torch.empty(n * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    ts = []
    for _ in range(10):
        mod.process_batch_device(bits.data_ptr(), n, out.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        ts.append(mod.kernel_times())
    tot = float(np.median([sum(t for _, t in tt) for tt in ts[3:]]))
    names = [k for k, _ in ts[0]]
    print("fir_kernel=%d" % fk, names, "total %.4f ms" % tot, flush=True)
    mod.close()
