#!/usr/bin/env python
"""Generates tests/golden/mufu_rcp_table.npy ON THE GPU BOX: MUFU.RCP((float)i) for i = 0..1023, read
back through the parity tap hr_debug_rcp_table. The CPU oracle's NVIDIA-OpenCL arithmetic
(oracle/hr_oracle.h, HRO_ARITH_NVCL) needs it to restate `x / y` = div.full.f32 = x * MUFU.RCP(y).

  gpurun -- 'python tests/golden/make_rcp_table.py'   ->  gpurun_out/mufu_rcp_table.npy  (copy it here)
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import hr_pkg

hr = hr_pkg.load()
t = hr.debug_rcp_table(1024)
ref = np.float32(1.0) / np.arange(1024, dtype=np.float32)
ulp = np.abs(t[1:].view(np.int32).astype(np.int64) - ref[1:].view(np.int32).astype(np.int64))
print("MUFU.RCP vs correctly rounded 1/i: max %d ulp, %d of 1023 differ; rcp(255) = %r (1/255 = %r)" % (ulp.max(), int((ulp > 0).sum()), t[255], ref[255]))
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
np.save(out / "mufu_rcp_table.npy", t)
