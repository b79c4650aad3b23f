#!/usr/bin/env python
"""C2 (TM I, 1024 TFs, FIR default taps): the two-kernel step (fir_kernel = 2) against the fused kernel (3)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dabmod_loader  # noqa: E402

dm = dabmod_loader.load()
for n in [int(a) for a in sys.argv[1:]] or [1024]:
  for fk in (2, 3, 2, 3):
      mod = dm.Modulator(mode=1, fir_taps="default", max_batch=n)
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
      med = {k: float(np.median([dict(t)[k] for t in ts[3:]])) for k, _ in ts[0]}
      print("n_tf=%d fir_kernel=%d" % (n, fk), med, "total %.4f ms" % sum(med.values()), flush=True)
      mod.close()
