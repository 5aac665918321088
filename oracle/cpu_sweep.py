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
        import numpy.linalg  # noqa: F401  (load every BLAS/LAPACK the solve can reach BEFORE limiting their thread pools)
        import scipy.linalg  # noqa: F401
        import scipy.linalg.cython_lapack  # noqa: F401  (numba's np.linalg.inv binds this one)
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
    proj = cfg.get("projectile", (1, 0))
    t0 = time.perf_counter()
    wss = []
    for (tgt, Elab, Rch, kin) in energies:
        wss.append((tgt, Elab, orc.ElasticWorkspace(N, lmax, Rch, kin, angles, tgt[1] * proj[1])))
    # untimed warm-up: numba compiles the njit core on the first call in every fresh process (the
    # reference's users pay this once per session, not per sample)
    with np.errstate(over="ignore"):
        c, so, co = orc.kduq_potentials(wss[0][2].rgrid, proj, wss[0][0], wss[0][1], rows[0])
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
                c, so, co = orc.kduq_potentials(ws.rgrid, proj, tgt, Elab, p)
            dsdo, Ay, Q, tot, rxn = ws.xs(c, so, co, count=count)
            acc += rxn
            npairs += 1
    dt = time.perf_counter() - t0
    return dict(solves=count[0], pairs=npairs, seconds=dt, setup=setup, checksum=acc)


def run(energy_table, rows, N=40, lmax=30, angles=None, procs=None, pairs_per_worker=2000, projectile=(1, 0)):
    """energy_table: list of (target(A,Z), Elab, Rch_fm, kin tuple).  Each worker takes a round-robin
    slice of the energies and runs ``pairs_per_worker`` (sample, energy) pairs over it."""
    procs = procs or os.cpu_count() or 1
    angles = np.linspace(0.1, np.pi, 180) if angles is None else angles
    cfg = dict(N=N, lmax=lmax, angles=angles, projectile=tuple(projectile))
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


# ---------------------------------------------------------------------------------------------------------------
# general solves (BASELINE configs 4 and 5): Solver.solve of the reference, one system per call (rmatrix.py:180-246)
# ---------------------------------------------------------------------------------------------------------------
def _work_general(job):
    from jitr_b200 import workloads
    from oracle import jitr_oracle as orc

    kind, wid, n_systems = job
    t0 = time.perf_counter()
    if kind == "config4":
        c = workloads.CONFIG4
        N, nch = c["N"], 1
        kin = orc.semi_relativistic_kinematics(workloads.TARGETS[1][2], workloads.MASS_N, c["Elab"], 0)
        _, Ecm, mu, k, eta = kin
        a = c["Rch"] * k
        x, _ = orc.legendre_mesh(N)
        r = x * a / k
        rg, rpg = np.meshgrid(r, r, indexing="ij")
        smp = workloads.config4_samples(4, seed=100 + wid)
        ls = [0, 5, 11, 23]
        pots = [workloads.perey_buck_kernel(rg, rpg, l, *smp[i]) for i, l in enumerate(ls)]
        asym = [tuple(np.array([v]) for v in orc.coulomb_hankel(l, eta, a)) for l in ls]
        args = [dict(l=[l], E=[Ecm], k=[k], mu=[mu], asym=asym[i], nonlocal_potential=pots[i]) for i, l in enumerate(ls)]
    else:
        c = workloads.CONFIG5
        N, nch = c["N"], 3
        a, k = c["a"], c["k"]
        x, _ = orc.legendre_mesh(N)
        r = x * a / k
        V = workloads.config5_potentials(r, 4, seed=200 + wid)
        ls = [0, 3, 6, 10]
        args = []
        for i, l in enumerate(ls):
            H = np.array([orc.coulomb_hankel(l, c["eta"], a) for _ in range(nch)]).T  # the example reuses k, eta for all channels
            args.append(dict(l=[l] * nch, E=list(c["E"]), k=[k] * nch, mu=[c["mu"]] * nch, asym=tuple(np.ascontiguousarray(H[j]) for j in range(4)),
                             local_potential=V[i]))
    # Solver.solve(..., free_matrix=F): the free matrix of a partial wave is precomputed once, as a careful user of the
    # reference does (rmatrix.py:206-214); per call: interaction matrix from the mesh values, A = F + V, inverse, R, S
    b = orc.boundary_vector(N)
    w = np.zeros(nch)
    w[0] = 1
    for g in args:
        g["F"] = orc.free_matrix(N, a, g["l"], g["E"], g["mu"])
    orc.solver_solve(N, a, **{k_: v for k_, v in args[0].items() if k_ != "F"})  # JIT warm-up, untimed
    setup = time.perf_counter() - t0
    acc = 0.0
    t0 = time.perf_counter()
    for i in range(n_systems):
        g = args[i % len(args)]
        Vm = orc.interaction_matrix(N, g["k"][0], g["E"][0], a, nch, g.get("local_potential"), g.get("nonlocal_potential"))
        R, S, _, up = orc.solve_smatrix(g["F"] + Vm, b, *g["asym"], w, a, nch, N)
        acc += float(abs(S[0, 0]))
    dt = time.perf_counter() - t0
    return dict(solves=n_systems, seconds=dt, setup=setup, checksum=acc)


def run_general(kind, procs=None, systems_per_worker=2000):
    """kind: "config4" (N = 60 nonlocal, n = 60) or "config5" (3 x 50 coupled, n = 150).  One process per core, each timing
    `systems_per_worker` calls of the oracle's Solver.solve on that configuration's systems."""
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_init, initargs=({},)) as pool:
        res = pool.map(_work_general, [(kind, w, systems_per_worker) for w in range(procs)])
    solves = sum(r["solves"] for r in res)
    tmax = max(r["seconds"] for r in res)
    return dict(solves_per_s=solves / tmax, solves=solves, seconds=tmax, setup_seconds=max(r["setup"] for r in res), procs=procs)
