"""Oracle side of the large-sample parity gate: the reference's user loop
(examples/notebooks/kduq_uq_demo.ipynb cell 16 -> xs/elastic.py:104-183, 332-381) restated by
oracle/jitr_oracle.py, evaluated for EVERY (parameter row, workspace) pair of a sweep on all host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see the header of oracle/jitr_oracle.py): used by tests/, by
tools/parity_config2.py and by bench.py's ``parity`` key as the checker, never as the thing measured.
"""
from __future__ import annotations

import multiprocessing as mp
import os

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
    from oracle import jitr_oracle as orc

    e_idx, (tgt, Elab, Rch, kin) = job
    cfg = _W["cfg"]
    N, lmax, angles, rows, proj = cfg["N"], cfg["lmax"], cfg["angles"], cfg["rows"], cfg["projectile"]
    Zz = tgt[1] * proj[1]
    ws = orc.ElasticWorkspace(N, lmax, Rch, kin, angles, Zz)
    B = len(rows)
    sp = np.zeros((B, lmax + 1), dtype=np.complex128)
    sm = np.zeros((B, lmax + 1), dtype=np.complex128)
    nl = np.zeros(B, dtype=np.int32)
    dsdo = np.zeros((B, len(angles)))
    tr = np.zeros((B, 2))
    for i, p in enumerate(rows):
        with np.errstate(over="ignore"):
            c, so, co = orc.kduq_potentials(ws.rgrid, proj, tgt, Elab, p)
        a, b = ws.smatrix(c, so, co)
        L = len(a)
        sp[i, :L], sm[i, :L], nl[i] = a, b, L
        d, _, _, tot, rxn = orc.differential_elastic_xs(ws.k, ws.angles, a, b, ws.P, ws.P1, ws.f_c, ws.sigma_l)
        dsdo[i], tr[i, 0], tr[i, 1] = d, tot, rxn
    return e_idx, sp, sm, nl, dsdo, tr


def run(energy_table, rows, N=40, lmax=30, angles=None, projectile=(1, 0), procs=None):
    """All rows x all workspaces.  Returns dict(splus, sminus [B][E][L1] c128, nl [B][E], dsdo [B][E][nth], tot_rxn [B][E][2])."""
    procs = procs or os.cpu_count() or 1
    angles = np.linspace(0.1, np.pi, 180) if angles is None else angles
    cfg = dict(N=N, lmax=lmax, angles=angles, rows=np.asarray(rows, dtype=np.float64), projectile=tuple(projectile))
    B, E = len(rows), len(energy_table)
    out = dict(splus=np.zeros((B, E, lmax + 1), dtype=np.complex128), sminus=np.zeros((B, E, lmax + 1), dtype=np.complex128),
               nl=np.zeros((B, E), dtype=np.int32), dsdo=np.zeros((B, E, len(angles))), tot_rxn=np.zeros((B, E, 2)))
    ctx = mp.get_context("fork")
    with ctx.Pool(min(procs, E), initializer=_init, initargs=(cfg,)) as pool:
        for e_idx, sp, sm, nl, dsdo, tr in pool.imap_unordered(_work, list(enumerate(energy_table))):
            out["splus"][:, e_idx], out["sminus"][:, e_idx], out["nl"][:, e_idx] = sp, sm, nl
            out["dsdo"][:, e_idx], out["tot_rxn"][:, e_idx] = dsdo, tr
    return out
