"""CPU baseline runner: the oracle's restatement of the reference's user loop
(examples/notebooks/kduq_uq_demo.ipynb cell 16: ``omp(rgrid, ...)`` -> ``DifferentialWorkspace.xs``)
on all host cores, one process per core, single-threaded BLAS.

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/jitr_oracle.py header).  Used by bench.py's
``cpu_baseline`` leg and by ``bench.py --impl reference``.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_W = {}


def _init(cfg):
    try:
        from threadpoolctl import threadpool_limits

        _W["limit"] = threadpool_limits(1)
    except Exception:
        pass
    _W["cfg"] = cfg


def _work(job):
    """job = (worker_id, list of (target, Elab, Rch, kin, masses), params rows, pairs_per_energy)."""
    from oracle import jitr_oracle as orc

    wid, energies, rows, reps = job
    cfg = _W["cfg"]
    N, lmax, angles = cfg["N"], cfg["lmax"], cfg["angles"]
    t0 = time.perf_counter()
    wss = []
    for (tgt, Elab, Rch, kin) in energies:
        wss.append((tgt, Elab, orc.ElasticWorkspace(N, lmax, Rch, kin, angles, 0)))
    # untimed warm-up: numba compiles the njit core on the first call in every fresh process (the
    # reference's users pay this once per session, not per sample)
    with np.errstate(over="ignore"):
        c, so, co = orc.kduq_potentials(wss[0][2].rgrid, (1, 0), wss[0][0], wss[0][1], rows[0])
    wss[0][2].xs(c, so, co)
    setup = time.perf_counter() - t0
    count = [0]
    npairs = 0
    acc = 0.0
    t0 = time.perf_counter()
    for (tgt, Elab, ws) in wss:
        for i in range(reps):
            p = rows[(i + 7 * wid) % len(rows)]
            with np.errstate(over="ignore"):
                c, so, co = orc.kduq_potentials(ws.rgrid, (1, 0), tgt, Elab, p)
            dsdo, Ay, Q, tot, rxn = ws.xs(c, so, co, count=count)
            acc += rxn
            npairs += 1
    dt = time.perf_counter() - t0
    return dict(solves=count[0], pairs=npairs, seconds=dt, setup=setup, checksum=acc)


def run(energy_table, rows, N=40, lmax=30, angles=None, procs=None, pairs_per_worker=2000):
    """energy_table: list of (target(A,Z), Elab, Rch_fm, kin tuple).  Each worker takes a round-robin
    slice of the energies and runs ``pairs_per_worker`` (sample, energy) pairs over it."""
    procs = procs or os.cpu_count() or 1
    angles = np.linspace(0.1, np.pi, 180) if angles is None else angles
    cfg = dict(N=N, lmax=lmax, angles=angles)
    jobs = []
    for w in range(procs):
        mine = [energy_table[i] for i in range(w, len(energy_table), procs)] or [energy_table[w % len(energy_table)]]
        # at most 4 workspaces per worker keeps the (untimed) mpmath setup short
        mine = mine[:4]
        reps = max(1, pairs_per_worker // len(mine))
        jobs.append((w, mine, rows, reps))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs, initializer=_init, initargs=(cfg,)) as pool:
        res = pool.map(_work, jobs)
    wall = time.perf_counter() - t0
    solves = sum(r["solves"] for r in res)
    pairs = sum(r["pairs"] for r in res)
    tmax = max(r["seconds"] for r in res)
    return dict(
        solves_per_s=solves / tmax, samples_per_s=pairs / tmax, solves=solves, pairs=pairs, seconds=tmax,
        setup_seconds=max(r["setup"] for r in res), wall_seconds=wall, procs=procs,
        solves_per_pair=solves / max(pairs, 1),
    )
