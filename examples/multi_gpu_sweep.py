#!/usr/bin/env python
"""The KDUQ sweep of elastic_kduq_sweep.py on several GPUs of one node: the sample axis is sharded over one process per GPU, nothing
is exchanged during the solve, and the angular distributions of every rank land in ONE tensor on rank 0 -- pushed there chunk by
chunk by the copy engines while the next chunk is being solved (jitr_b200.sharding.PeerWindow).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/multi_gpu_sweep.py [samples_per_gpu]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import jitr_b200  # noqa: E402
from jitr_b200.optical_potentials import kduq  # noqa: E402
from jitr_b200.reactions import ElasticReaction  # noqa: E402
from jitr_b200.sharding import PeerWindow, shard_bounds  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
n_samples = per_gpu * world
rx = ElasticReaction((208, 82), (1, 0))
angles = np.linspace(0.1, np.pi, 180)
solver = jitr_b200.Solver(40)
sweep = jitr_b200.ElasticSweep([jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(E), 12.0, solver, 30, angles)
                                for E in (10.0, 30.0, 60.0)], dev)
model = kduq.KDUQ((1, 0))

# every rank draws ITS block of the synthetic posterior (seeded by the global sample index range it owns)
lo, hi = shard_bounds(n_samples, world, rank)
central = kduq.get_kd03((1, 0))
rng = np.random.default_rng(1000 + rank)
rows = np.tile(central, (hi - lo, 1)) * (1 + 0.02 * rng.standard_normal((hi - lo, central.size)))
rows[:, -2:] = central[-2:]
params = torch.zeros((hi - lo, 40), dtype=torch.float64, device=dev)
params[:, : rows.shape[1]] = torch.tensor(rows, device=dev)

win = PeerWindow((n_samples, sweep.nE, len(angles)), torch.float64, dev) if world > 1 else None
dsdo = torch.empty((hi - lo, sweep.nE, len(angles)), dtype=torch.float64, device=dev)
n_chunks = 4
for c in range(n_chunks):
    c0, c1 = c * (hi - lo) // n_chunks, (c + 1) * (hi - lo) // n_chunks
    S, nl, status = sweep.smatrix_params(model, params[c0:c1])
    assert int(status.abs().sum()) == 0
    sweep.observables(S, nl, want_Ay=False, want_Q=False,
                      out=(dsdo[c0:c1], None, None, torch.empty((c1 - c0, sweep.nE, 2), dtype=torch.float64, device=dev)))
    if win is not None:
        win.push(dsdo[c0:c1], lo + c0)   # copy engines over NVLink, beside the next chunk's kernels
if win is not None:
    win.flush()                          # + one 4-byte all_reduce: every rank's rows have landed
torch.cuda.synchronize()

if rank == 0:
    full = win.buf if win is not None else dsdo
    band = jitr_b200.uq.percentiles(full, [16, 50, 84]).cpu().numpy()   # (3, E, angles) from all n_samples on rank 0
    print(f"{n_samples} samples on {world} GPU(s); dsigma/dOmega at 30 MeV, 90 deg: {band[1, 1, 89]:.4e} "
          f"[{band[0, 1, 89]:.4e}, {band[2, 1, 89]:.4e}] mb/sr")
if win is not None:
    win.close()
    dist.destroy_process_group()
