#!/usr/bin/env python
"""Device reductions over the sample axis: percentile bands (jitr_b200_percentiles) on a sweep-sized block."""
import json
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from bench_general import timeit  # noqa: E402

from jitr_b200 import uq  # noqa: E402

S, M = 100000, 64 * 180 // 4  # a quarter of configs[1]'s (energies x angles) columns: 2.3 GB
x = torch.rand((S, M), dtype=torch.float64, device="cuda")
ms = timeit(lambda: uq.percentiles(x, [16, 84]))
print(json.dumps({"samples": S, "columns": M, "quantiles": 2, "ms": ms, "GBps_of_input_per_pass": 9 * S * M * 8 / ms / 1e6,
                  "input_GB": S * M * 8 / 1e9}))
