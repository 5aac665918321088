#!/usr/bin/env python
"""Sweep kernel alone on config 2 (S samples x 64 workspaces), for A/B runs of tuning variants:
    JITR_B200_LIB=build/variants/<name>/libjitr_b200.so python tools/bench_sweep_quick.py [samples] [config]
config 2 (default): N = 40 KDUQ neutrons; config 3: p+48Ca N = 60 lmax = 40 on the 416-row proton posterior tiled."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jitr_b200  # noqa: E402
from jitr_b200 import workloads  # noqa: E402
from jitr_b200.optical_potentials import kduq  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if cfg == 2:
    sweep = jitr_b200.ElasticSweep(workloads.config2_workspaces())
    model = kduq.KDUQ((1, 0))
    rows = workloads.synthetic_kduq_samples(S)
    n = 40
else:
    rx, ws = workloads.config3_workspace()
    sweep = jitr_b200.ElasticSweep([ws])
    model = kduq.KDUQ((1, 1))
    rows = workloads.synthetic_kduq_samples(S, projectile=(1, 1))
    n = 60
p = np.zeros((len(rows), 40))
p[:, : rows.shape[1]] = rows
pt = torch.tensor(p, device="cuda")
out = sweep._alloc_S(len(rows))
for _ in range(2):
    sweep.smatrix_params(model, pt, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
a.record()
for _ in range(reps):
    sweep.smatrix_params(model, pt, out=out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
solves = int((2 * out[1].to(torch.int64) - 1).sum())
print(json.dumps({"lib": os.environ.get("JITR_B200_LIB", "default"), "config": cfg, "samples": S, "ms": ms, "solves_per_s": solves / ms * 1e3,
                  "alg_tflops": solves * workloads.flops_per_solve(n) / ms / 1e9, "status_bad": int((out[2] != 0).sum())}))
