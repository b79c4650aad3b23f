#!/usr/bin/env python
"""Where a resampler kernel's time goes: runs C3 / C5 with the "res_dbg" knob (parts of the kernel skipped, results
wrong) and prints the kernel time of each variant.  usage: python tools/res_dbg.py [c3|c5] [masks...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dabmod_loader  # noqa: E402
import bench  # noqa: E402

dm = dabmod_loader.load()
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
masks = [int(x) for x in sys.argv[2:]] or [0, 1, 2, 3, 4, 8, 12, 15]
RK = os.environ.get("RES_KERNEL", "1")
for name, kw, ntf in bench.other_configs():
    if not name.startswith(which + " "):
        continue
    for poly in (True, False):
        kw2 = dict(kw)
        if not poly:
            kw2.pop("poly", None)
        mod = dm.Modulator(max_batch=ntf, **kw2)
        bits = torch.randint(0, 256, (ntf, mod.tf_in_bytes), dtype=torch.uint8).cuda()
        out = torch.empty(ntf * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
        st = torch.cuda.Stream()
        mod.set_param("profile", 1)
        mod.set_param("res_kernel", RK)
        for m in masks:
            mod.set_param("res_dbg", m)
            ts = []
            for _ in range(6):
                mod.process_batch_device(bits.data_ptr(), ntf, out.data_ptr(), st.cuda_stream)
                torch.cuda.synchronize()
                ts.append([t for k, t in mod.kernel_times() if k.startswith("k_resample")][0])
            print("%s rk=%s poly=%d dbg=%2d  %.4f ms" % (which, RK, poly, m, float(np.median(ts[2:]))), flush=True)
        mod.close()
