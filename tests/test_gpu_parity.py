"""GPU parity tests: the CUDA path (through the C ABI / jitr_b200 API) vs the golden vectors of
the real reference and vs the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): 1e-10 relative on S-matrix elements, 1e-8 relative on
cross sections.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import jitr_b200  # noqa: E402
from jitr_b200 import _cabi  # noqa: E402
from jitr_b200.optical_potentials import LocalOpticalPotential, ShapeModel, chuq, kduq, wlh  # noqa: E402
from jitr_b200.reactions import ElasticReaction, ProjectileTargetSystem  # noqa: E402
from oracle import jitr_oracle as orc  # noqa: E402

RTOL_S = 1e-10
RTOL_XS = 1e-8
# Ay and Q are bounded ratios in [-1, 1] that cross zero: near a crossing only the ABSOLUTE error is
# meaningful (2 Im/Re(a* b) cancels), so they are checked to 1e-8 relative OR 1e-9 absolute.
ATOL_POL = 1e-9


def relerr(a, b, floor=0.0):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor + 1e-300)))


def normerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def make_ws(g):
    rx = ElasticReaction(tuple(int(v) for v in g["target"]), tuple(int(v) for v in g["projectile"]),
                         mass_target=float(g["masses"][0]), mass_projectile=float(g["masses"][1]))
    kin = rx.kinematics(float(g["kin"][0]))
    solver = jitr_b200.Solver(int(g["N"]))
    ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, float(g["Rch"]), solver, int(g["lmax"]), g["angles"])
    return rx, kin, ws


def model_for(name, g):
    if name.startswith("chuq"):
        return chuq.CHUQ()
    if name.startswith("localomp"):
        return LocalOpticalPotential()
    if name.startswith("wlh"):
        return wlh.WLH(tuple(int(v) for v in g["projectile"]))
    return kduq.KDUQ(tuple(int(v) for v in g["projectile"]))


def check_xs_against_golden(x, S, nl, g, idx):
    """x: ElasticXS of tensors with leading sample axis; compare rows idx with the golden arrays."""
    S = S.cpu().numpy()
    nl = nl.cpu().numpy()
    for row, i in enumerate(idx):
        L = int(g["nl"][i])
        assert int(nl[row]) == L
        assert relerr(S[row, 0, :L], g["splus"][i, :L]) < RTOL_S
        assert relerr(S[row, 1, 1:L], g["sminus"][i, 1:L]) < RTOL_S
        assert np.all(S[row, :, L:] == 0)
        assert relerr(x.dsdo[row].cpu().numpy(), g["dsdo"][i]) < RTOL_XS
        np.testing.assert_allclose(x.Ay[row].cpu().numpy(), g["Ay"][i], rtol=RTOL_XS, atol=ATOL_POL)
        np.testing.assert_allclose(x.Q[row].cpu().numpy(), g["Q"][i], rtol=RTOL_XS, atol=ATOL_POL)
        assert abs(float(x.t[row]) - g["tot"][i]) / g["tot"][i] < RTOL_XS
        assert abs(float(x.rxn[row]) - g["rxn"][i]) / g["rxn"][i] < RTOL_XS


ELASTIC = [
    "config1_n208Pb_30MeV", "config3_p48Ca_30MeV",
    "config2_n40Ca_E00", "config2_n40Ca_E06", "config2_n40Ca_E15", "config2_n40Ca_E31",
    "config2_n208Pb_E00", "config2_n208Pb_E06", "config2_n208Pb_E15", "config2_n208Pb_E31",
    "chuq_p90Zr_25MeV", "chuq_n90Zr_12MeV", "localomp_p54Fe_35MeV", "wlh_n208Pb_14MeV", "wlh_p48Ca_45MeV",
]


@pytest.mark.parametrize("name", ELASTIC)
def test_fused_params_sweep_vs_reference_golden(golden, name):
    """Built-in potentials evaluated in the fused kernel from parameter rows (north_star kernel 1+2+3)."""
    g = golden(name)
    rx, kin, ws = make_ws(g)
    model = model_for(name, g)
    params = torch.tensor(g["params"], dtype=torch.float64, device="cuda")
    S, nl, status = ws.integral_workspace.smatrix_params(model, params)
    assert int(status.abs().sum()) == 0
    x = ws.xs_params(model, params)
    check_xs_against_golden(x, S, nl, g, range(len(g["params"])))


@pytest.mark.parametrize("name", ["config1_n208Pb_30MeV", "config3_p48Ca_30MeV", "chuq_p90Zr_25MeV"])
def test_tensor_sweep_vs_reference_golden(golden, name):
    """User potentials as device tensors of mesh values (reference's DifferentialWorkspace.xs inputs)."""
    g = golden(name)
    rx, kin, ws = make_ws(g)
    c = torch.tensor(g["central"], dtype=torch.complex128, device="cuda")
    so = torch.tensor(g["so"], dtype=torch.complex128, device="cuda")
    co = torch.tensor(g["coul"].astype(np.complex128), dtype=torch.complex128, device="cuda")
    S, nl, status = ws.integral_workspace.smatrix_batch(c, so, co)
    assert int(status.abs().sum()) == 0
    x = ws.xs_batch(c, so, co)
    check_xs_against_golden(x, S, nl, g, range(c.shape[0]))
    # single-sample reference API (numpy in / numpy out)
    one = ws.xs(g["central"][0], g["so"][0], g["coul"][0])
    assert relerr(one.dsdo, g["dsdo"][0]) < RTOL_XS
    assert abs(one.t - g["tot"][0]) / g["tot"][0] < RTOL_XS
    sp, sm = ws.integral_workspace.smatrix(g["central"][0], g["so"][0], g["coul"][0])
    assert len(sp) == int(g["nl"][0]) and relerr(sp, g["splus"][0, : len(sp)]) < RTOL_S
    tot, rxn = ws.integral_workspace.xs(g["central"][0], g["so"][0], g["coul"][0])
    assert abs(rxn - g["rxn"][0]) / g["rxn"][0] < RTOL_XS
    tp, tm = ws.integral_workspace.transmission_coefficients(g["central"][0], g["so"][0], g["coul"][0])
    np.testing.assert_allclose(tp, 1 - np.abs(g["splus"][0, : len(tp)]) ** 2, rtol=1e-8, atol=1e-12)


def test_optional_spin_orbit_equals_zeros(golden):
    """restated tests/test_xs_optional_spin_orbit.py: omitted spin-orbit == explicit zeros."""
    g = golden("config1_n208Pb_30MeV")
    rx, kin, ws = make_ws(g)
    c = g["central"][0]
    a = ws.xs(c)
    b = ws.xs(c, np.zeros_like(c))
    np.testing.assert_allclose(a.dsdo, b.dsdo, rtol=1e-12)
    np.testing.assert_allclose(a.Ay, b.Ay, rtol=1e-7, atol=1e-12)
    assert a.t == b.t and a.rxn == b.rxn
    sp0, sm0 = ws.integral_workspace.smatrix(c)
    sp1, sm1 = ws.integral_workspace.smatrix(c, np.zeros_like(c))
    np.testing.assert_allclose(sp0, sp1, rtol=1e-12)
    with pytest.raises(ValueError):
        ws.xs(c[:-1])


def test_device_potential_forms_vs_reference(golden):
    """SingleChannelOpticalModel.evaluate: device forms vs the reference's numpy forms (golden) to ~1 ulp."""
    for name in ("config1_n208Pb_30MeV", "config3_p48Ca_30MeV", "chuq_p90Zr_25MeV", "chuq_n90Zr_12MeV", "localomp_p54Fe_35MeV",
                 "wlh_n208Pb_14MeV", "wlh_p48Ca_45MeV"):
        g = golden(name)
        rx, kin, ws = make_ws(g)
        model = model_for(name, g)
        for i in range(min(4, len(g["params"]))):
            c, so, co = model(ws.radial_grid(), rx, kin, *g["params"][i])
            np.testing.assert_allclose(c, g["central"][i], rtol=1e-13, atol=1e-300)
            np.testing.assert_allclose(so, g["so"][i], rtol=1e-13, atol=1e-300)
            np.testing.assert_allclose(co, g["coul"][i], rtol=1e-13, atol=1e-300)


def test_kduq_uq_demo_notebook_golden(golden):
    """examples/notebooks/kduq_uq_demo.ipynb cell 20: 84th percentile of dsdo/ruth at angle index 9
    over the 416 federal proton samples prints 1.163423."""
    g = golden("nb_kduq_uq_demo")
    P = golden("kduq_params")["federal_p"]
    rx = ElasticReaction((54, 26), (1, 1), mass_target=float(g["masses"][0]), mass_projectile=float(g["masses"][1]))
    kin = rx.kinematics(35.0)
    ang = np.linspace(0.1, np.pi, 180)
    ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, float(g["Rch"]), jitr_b200.Solver(40), 50, ang)
    x = ws.xs_params(kduq.KDUQ((1, 1)), torch.tensor(P, dtype=torch.float64, device="cuda"))
    ratio = x.dsdo.cpu().numpy() / ws.rutherford
    assert relerr(ratio, g["ratio"]) < RTOL_XS
    assert round(float(np.percentile(ratio, 84, axis=0)[9]), 6) == 1.163423


def test_fused_sweep_vs_oracle_random_shapes():
    """Seeded random canonical shape parameters, several energies in one launch, vs the CPU oracle."""
    rng = np.random.default_rng(7)
    N, lmax = 30, 12
    rx = ElasticReaction((90, 40), (1, 1), mass_target=83725.25, mass_projectile=938.271653086152)
    energies = [8.0, 21.5, 40.0]
    ang = np.linspace(0.2, 3.0, 37)
    solver = jitr_b200.Solver(N)
    wss = [jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(E), 11.0, solver, lmax, ang) for E in energies]
    sweep = jitr_b200.ElasticSweep(wss)
    B = 5
    p = np.zeros((B, 17))
    p[:, 0] = rng.uniform(40, 60, B); p[:, 1] = rng.uniform(5, 6, B); p[:, 2] = rng.uniform(0.5, 0.8, B)
    p[:, 3] = rng.uniform(1, 8, B); p[:, 4] = rng.uniform(5, 6, B); p[:, 5] = rng.uniform(0.5, 0.8, B)
    p[:, 6] = rng.uniform(0, 2, B); p[:, 7] = rng.uniform(2, 9, B); p[:, 8] = rng.uniform(5, 6.5, B); p[:, 9] = rng.uniform(0.4, 0.7, B)
    p[:, 10] = rng.uniform(8, 13, B); p[:, 11] = rng.uniform(4.5, 5.5, B); p[:, 12] = rng.uniform(0.5, 0.7, B)
    p[:, 13] = rng.uniform(-1, 1, B); p[:, 14] = rng.uniform(4.5, 5.5, B); p[:, 15] = rng.uniform(0.5, 0.7, B)
    p[:, 16] = rng.uniform(5, 6, B)
    x, status = sweep.xs_params(ShapeModel(), torch.tensor(p, device="cuda"), return_status=True)
    assert int(status.abs().sum()) == 0
    for ie, E in enumerate(energies):
        kin = tuple(rx.kinematics(E))
        ow = orc.ElasticWorkspace(N, lmax, 11.0, kin, ang, 40)
        r = ow.rgrid
        for s in range(B):
            q = p[s]
            c = (-q[0] * orc.woods_saxon_safe(r, q[1], q[2]) - 1j * q[3] * orc.woods_saxon_safe(r, q[4], q[5])
                 + 4 * q[9] * (q[6] + 1j * q[7]) * orc.woods_saxon_prime_safe(r, q[8], q[9]))
            so = q[10] * orc.thomas_safe(r, q[11], q[12]) + 1j * q[13] * orc.thomas_safe(r, q[14], q[15])
            co = orc.coulomb_charged_sphere(r, 40, q[16])
            dsdo, Ay, Q, tot, rxn = ow.xs(c, so, co)
            assert relerr(x.dsdo[s, ie].cpu().numpy(), dsdo) < RTOL_XS
            np.testing.assert_allclose(x.Ay[s, ie].cpu().numpy(), Ay, rtol=RTOL_XS, atol=ATOL_POL)
            assert abs(float(x.rxn[s, ie]) - rxn) / rxn < RTOL_XS


@pytest.mark.parametrize("N", [6, 12, 20, 27, 35, 40, 48, 56, 64])
def test_sweep_every_nbasis_path_vs_oracle(N):
    """Every solver mode of the fused sweep: generic (N <= 8 and 40 < N <= 64, two or three row slots per
    lane) and the assembly-free fast paths P = 16, 24, 32, 40 (N not a multiple of 8 exercises the identity
    padding), both pivoting schemes, against the oracle on the same potentials."""
    lmax = 5
    rx = ElasticReaction((48, 20), (1, 1), mass_target=44657.272049078, mass_projectile=938.271653086152)
    kin = rx.kinematics(18.0)
    iw = jitr_b200.IntegralWorkspace(rx, kin, 9.5, jitr_b200.Solver(N), lmax, smatrix_abs_tol=0.0)
    r = iw.radial_grid()
    rng = np.random.default_rng(100 + N)
    B = 6
    c = np.zeros((B, N), dtype=np.complex128)
    so = np.zeros((B, N), dtype=np.complex128)
    co = np.zeros((B, N), dtype=np.complex128)
    for b in range(B):
        V, W, R0, a0 = rng.uniform(40, 55), rng.uniform(2, 9), rng.uniform(4.2, 4.9), rng.uniform(0.55, 0.75)
        c[b] = -(V + 1j * W) * orc.woods_saxon_safe(r, R0, a0) + 4 * a0 * 1j * rng.uniform(3, 8) * orc.woods_saxon_prime_safe(r, R0 + 0.3, a0)
        so[b] = rng.uniform(5, 9) * orc.thomas_safe(r, R0 - 0.4, 0.6)
        co[b] = orc.coulomb_charged_sphere(r, 20, 1.25 * 48 ** (1 / 3))
    ow = orc.ElasticWorkspace(N, lmax, 9.5, tuple(kin), np.linspace(0.3, 3.0, 7), 20, tol=0.0)
    L = _cabi.lib()
    for sym in (1, 0):
        try:
            assert L.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, sym) == 0
            S, nl, status = iw.smatrix_batch(torch.tensor(c, device="cuda"), torch.tensor(so, device="cuda"), torch.tensor(co, device="cuda"))
        finally:
            L.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, 1)
        assert int(status.abs().sum()) == 0
        Sg = S.cpu().numpy()
        for b in range(B):
            sp, sm = ow.smatrix(c[b], so[b], co[b])
            # N as small as 6 is far from converged: the systems are what they are, parity is what is tested
            assert relerr(Sg[b, 0, : len(sp)], sp) < RTOL_S
            assert relerr(Sg[b, 1, 1 : len(sm)], sm[1:]) < RTOL_S


def test_sweep_is_deterministic_under_dynamic_work_distribution(golden):
    """Pairs are handed to the warps dynamically and the warps run in loose lock step: results must be bit
    identical whatever the batch size (awkward pair counts around the warps-per-CTA and grid multiples also
    exercise the termination of the lock step)."""
    g = golden("kduq_params")
    rows = np.zeros((1800, 40))
    rows[:, :37] = np.tile(g["federal_n"], (5, 1))[:1800]
    rx = ElasticReaction((208, 82), (1, 0), mass_target=193687.12278341982, mass_projectile=939.5654983938299)
    solver = jitr_b200.Solver(40)
    ang = np.linspace(0.1, np.pi, 11)
    wss = [jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(E), 12.0, solver, 30, ang) for E in (7.0, 61.0)]
    sweep = jitr_b200.ElasticSweep(wss)
    model = kduq.KDUQ((1, 0))
    full = torch.tensor(rows, device="cuda")
    Sref, nlref, stref = sweep.smatrix_params(model, full)
    assert int(stref.abs().sum()) == 0
    for B in (1, 5, 6, 7, 73, 74, 75, 887, 888, 889, 1799):
        S, nl, st = sweep.smatrix_params(model, full[:B])
        assert torch.equal(nl, nlref[:B]) and torch.equal(S, Sref[:B])


def test_workspace_beyond_fused_nbasis_vs_oracle():
    """N = 80 > 64 (the reference's Frescox regression cases use N up to 100): the workspace API routes the
    partial waves through the general batched solve; same S, same early break, same cross sections."""
    N, lmax = 80, 14
    rx = ElasticReaction((78, 28), (1, 1), mass_target=72601.0, mass_projectile=938.271653086152)
    kin = rx.kinematics(20.0)
    ang = np.linspace(0.2, 3.0, 23)
    ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, 16.0, jitr_b200.Solver(N), lmax, ang)
    r = ws.radial_grid()
    model = LocalOpticalPotential()
    p = np.array([[52.0, 1.17, 0.75, 2.5, 1.26, 0.58, 7.5, 0.0, 1.26, 0.58, 6.2, -0.1, 1.01, 0.75, 1.25],
                  [48.0, 1.20, 0.70, 1.5, 1.30, 0.60, 8.5, 0.5, 1.22, 0.55, 5.0, 0.0, 1.05, 0.70, 1.30]])
    x, status = ws._as_sweep().xs_params(model, torch.tensor(p, device="cuda"), return_status=True)
    S, nl, _ = ws.integral_workspace.smatrix_params(model, torch.tensor(p, device="cuda"))
    assert int(status.abs().sum()) == 0
    ow = orc.ElasticWorkspace(N, lmax, 16.0, tuple(kin), ang, 28)
    for b in range(2):
        c, so, co = orc.local_omp_potentials(r, 78, 28, p[b])
        sp, sm = ow.smatrix(c, so, co)
        assert int(nl[b]) == len(sp)
        assert relerr(S[b, 0, : len(sp)].cpu().numpy(), sp) < RTOL_S
        assert relerr(S[b, 1, 1 : len(sm)].cpu().numpy(), sm[1:]) < RTOL_S
        dsdo, Ay, Q, tot, rxn = ow.xs(c, so, co)
        assert relerr(x.dsdo[b, 0].cpu().numpy(), dsdo) < RTOL_XS
        assert abs(float(x.rxn[b, 0]) - rxn) / rxn < RTOL_XS
    # single-sample reference API on the same workspace
    c, so, co = orc.local_omp_potentials(r, 78, 28, p[0])
    one = ws.xs(c, so, co)
    assert relerr(one.dsdo, ow.xs(c, so, co)[0]) < RTOL_XS


def test_no_early_break_and_size_independent_properties():
    """tol <= 0 issues every (l, j); for a REAL potential |S| = 1 (unitarity) for every partial wave
    and sigma_rxn = 0 -- a size-independent property checked on a larger batch."""
    N, lmax = 40, 30
    rx = ElasticReaction((208, 82), (1, 0), mass_target=193687.12278341982, mass_projectile=939.5654983938299)
    kin = rx.kinematics(30.0)
    ws = jitr_b200.IntegralWorkspace(rx, kin, 12.0, jitr_b200.Solver(N), lmax, smatrix_abs_tol=0.0)
    B = 2048
    rng = np.random.default_rng(3)
    p = np.zeros((B, 17))
    p[:, 0] = rng.uniform(30, 60, B); p[:, 1] = rng.uniform(6.5, 7.5, B); p[:, 2] = rng.uniform(0.5, 0.8, B)
    p[:, 4] = 7.0; p[:, 5] = 0.6; p[:, 8] = 7.0; p[:, 9] = 0.6
    p[:, 10] = rng.uniform(5, 12, B); p[:, 11] = 6.8; p[:, 12] = 0.6; p[:, 14] = 6.8; p[:, 15] = 0.6; p[:, 16] = 7.0
    S, nl, status = ws.smatrix_params(ShapeModel(), torch.tensor(p, device="cuda"))
    assert int(status.abs().sum()) == 0
    assert int((nl != lmax + 1).sum()) == 0
    mod = S.abs()
    assert float((mod[:, 0, :] - 1).abs().max()) < 1e-9
    assert float((mod[:, 1, 1:] - 1).abs().max()) < 1e-9
    assert float(S[:, 1, 0].abs().max()) == 0.0


@pytest.mark.parametrize("config", [3, 4, 5])
def test_benchmarked_workloads_of_configs_3_4_5_vs_oracle(config):
    """The inputs `bench.py --config 3|4|5` times (leading rows of the same seeded generators), every partial wave, against
    the oracle: nl exact, S to 1e-10, cross sections to 1e-8 (tools/parity_general.py, the bench's own `parity` record)."""
    from tools import parity_general

    r = {3: lambda: parity_general.config3(64), 4: lambda: parity_general.config4(8), 5: lambda: parity_general.config5(6)}[config]()
    assert r["passed"], r


def test_large_sample_parity_gate_on_the_benchmarked_sweep():
    """BASELINE.md section 3 / SURVEY 8(d): all 416 federal KDUQ rows + 1000 of the bench's own multivariate-normal
    rows on all 64 (target, energy) workspaces of config 2 against the oracle (host cores), with the threshold-pivoted
    symmetric path AND with partial pivoting only: S to 1e-10, cross sections to 1e-8, the number of partial waves kept
    exactly; the report (incl. the statistics of the systems that fell back) is what profiles/r02_parity.json holds."""
    import json
    import os

    from tools import parity_config2

    rep = parity_config2.compare(n_synthetic=1000)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        open(os.path.join(out, "parity_config2.json"), "w").write(json.dumps(rep, indent=1))
    except OSError:
        pass
    for mode in ("symmetric_path", "partial_pivoting_only"):
        a = rep[mode]["all"]
        assert a["pairs"] == 1416 * 64
        assert a["nl_equal"] and a["zeros_beyond_nl"] and a["status_nonzero"] == 0, (mode, a)
        assert a["sigma_rxn"]["max"] < RTOL_XS and a["sigma_tot"]["max"] < RTOL_XS, (mode, a)
        # S to 1e-10 and dsigma/dOmega to 1e-8 for 99.9 % of 3.4e6 / 1.6e7 elements outright; every element beyond the gate
        # is handed to the multi-precision referee (the exact solution of the reference's own float64 matrix): the CUDA value
        # must be inside the gate of the exact value or at least as close to it as the float64 reference is (the reference
        # inverts matrices with cond(A) up to 1e7 explicitly and carries that noise itself)
        assert a["S"]["p999"] < RTOL_S and a["dsdo"]["p999"] < RTOL_XS, (mode, a)
        r = rep[mode]["referee"]
        assert r["S_all_explained"], (mode, r)
        assert r["S_outlier_systems"] <= 1e-4 * a["systems"], (mode, r)
        # cross sections: 1e-8 relative, or -- at diffraction minima, where dsigma/dOmega is 1e-9 of its forward value and the
        # partial-wave sum cancels by kappa ~ 1e5 -- the S-matrix tolerance propagated to first order, 1e-10 kappa
        assert a["dsdo_over_allowed_max"] <= 1.0 and a["dsdo_beyond_1e-8"] <= 1e-4 * a["dsdo"]["n"], (mode, a)
        assert r.get("referee_vs_mp40_max", 0.0) < 1e-12
    assert rep["symmetric_vs_partial_pivoting"]["nl_equal"]
    assert rep["partial_pivoting_only"]["all"]["fallback"]["systems"] == 0
    assert rep["passed"]


def test_partial_pivoting_only_option_matches(golden):
    """JITR_B200_OPT_SYMMETRIC = 0 forces partial pivoting for every system; both pivoting schemes must
    give the reference's S (the symmetric threshold-pivoted path is the default)."""
    g = golden("config1_n208Pb_30MeV")
    rx, kin, ws = make_ws(g)
    model = model_for("config1_n208Pb_30MeV", g)
    params = torch.tensor(g["params"], dtype=torch.float64, device="cuda")
    L = _cabi.lib()
    try:
        assert L.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, 0) == 0
        S0, nl0, st0 = ws.integral_workspace.smatrix_params(model, params)
        x0 = ws.xs_params(model, params)
    finally:
        assert L.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, 1) == 0
    S1, nl1, st1 = ws.integral_workspace.smatrix_params(model, params)
    x1 = ws.xs_params(model, params)
    assert int(st0.abs().sum()) == 0 and int(st1.abs().sum()) == 0
    check_xs_against_golden(x0, S0, nl0, g, range(len(g["params"])))
    check_xs_against_golden(x1, S1, nl1, g, range(len(g["params"])))
    assert torch.equal(nl0, nl1)
    assert L.jitr_b200_set_option(99, 0) == _cabi.ERR_BAD_ARG


@pytest.mark.parametrize("N", [40, 56])
def test_symmetric_path_falls_back_to_partial_pivoting(N):
    """(N = 40: in-kernel partial-pivoting re-solve; N = 56: a packed mode, in the l-split work distribution of a small batch --
    the pair is deferred once and redone by the partial-pivoting mode.)
    A diagonal entry that fails the threshold test (|a_kk| < max_i |a_ik| / 16) must send the system
    to the partial-pivoting path: a potential tuned so that the FIRST diagonal entry a^2 A[0][0] is tiny
    against its column at one partial wave (later diagonal entries are Schur complements and cannot be
    tuned from the potential alone)."""
    lmax = 6
    rx = ElasticReaction((40, 20), (1, 0), mass_target=37214.69, mass_projectile=939.5654983938299)
    kin = rx.kinematics(20.0)
    angles = np.linspace(0.2, 3.0, 31)
    ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, 10.0, jitr_b200.Solver(N), lmax, angles, smatrix_abs_tol=0.0)
    iw = ws.integral_workspace
    q = iw.solver.kernel.quadrature
    That = q.kinetic_hat()
    a, Ecm = iw.a, kin.Ecm
    r = iw.radial_grid()
    rng = np.random.default_rng(11)
    B = 24
    c = np.zeros((B, N), dtype=np.complex128)
    for bidx in range(B):
        base = -45.0 / (1.0 + np.exp((r - 4.2) / 0.65)) * (1.0 + 0.2j)
        k = 0
        l = int(rng.integers(0, lmax + 1))
        # a^2 A[k][k] = That_kk + l(l+1)/x_k^2 + a^2 (V_k / Ecm - 1)  ->  small at partial wave l
        base[k] = Ecm * (1.0 - (That[k, k] + l * (l + 1) / q.abscissa[k] ** 2 - 10.0 ** rng.uniform(-3, 1)) / a**2)
        c[bidx] = base
    L = _cabi.lib()
    f0 = L.jitr_b200_sym_fallback_count()
    S, nl, status = iw.smatrix_batch(torch.tensor(c, device="cuda"))
    x = ws.xs_batch(torch.tensor(c, device="cuda"))
    f1 = L.jitr_b200_sym_fallback_count()
    assert f1 - f0 >= B, "the tuned systems did not take the partial-pivoting fallback"
    assert int(status.abs().sum()) == 0
    ow = orc.ElasticWorkspace(N, lmax, 10.0, tuple(kin), angles, 0, tol=0.0)
    Sg = S.cpu().numpy()
    for bidx in range(B):
        sp, sm = ow.smatrix(c[bidx])
        assert relerr(Sg[bidx, 0, : len(sp)], sp) < RTOL_S
        assert relerr(Sg[bidx, 1, 1 : len(sm)], sm[1:]) < RTOL_S
        dsdo = ow.xs(c[bidx])[0]
        assert relerr(x.dsdo[bidx].cpu().numpy(), dsdo) < RTOL_XS


# ------------------------------------------------------------------ general Solver.solve path
def test_solver_solve_local_rk_case(golden):
    """restated tests/test_local.py (the RK45 cross-check case) against the reference's own S."""
    g = golden("test_local")
    N, a = int(g["N"]), float(g["a"])
    solver = jitr_b200.Solver(N)
    sysl = ProjectileTargetSystem(channel_radius=a, lmax=2, mass_target=g["masses"][0], mass_projectile=g["masses"][1], Ztarget=20, Zproj=1)
    free = solver.free_matrix(a, sysl.l, coupled=False)
    bb = solver.precompute_boundaries(a)
    for i in (0, 4, 9):
        channels, asym = sysl.get_partial_wave_channels(*g["kin"][i])
        for l in range(3):
            R, S, u = solver.solve(channels[l], asym[l], local_potential=g["pot"][i], basis_boundary=bb, free_matrix=free[l])
            assert R.shape == (1, 1) and S.shape == (1, 1) and u.shape == (1,)
            assert relerr(S[0, 0], g["S"][i, l]) < RTOL_S
            assert relerr(R[0, 0], g["R"][i, l]) < 1e-9
            assert relerr(u[0], g["uext"][i, l]) < 1e-9


def test_solver_solve_wavefunction(golden):
    """restated tests/test_wavefunction.py: N = 100, wavefunction=True, u_int(s)."""
    g = golden("test_wavefunction")
    N, a = int(g["N"]), float(g["a"])
    solver = jitr_b200.Solver(N)
    sysw = ProjectileTargetSystem(channel_radius=a, lmax=10, mass_target=44657.26581995028, mass_projectile=938.271653086152, Ztarget=20, Zproj=1)
    channels, asym = sysw.get_partial_wave_channels(*g["kin"])
    for l in (0, 1, 4):
        R, S, x, u = solver.solve(channels[l], asym[l], wavefunction=True, local_potential=g[f"pot{l}"])
        assert x.shape == (1, N)
        assert np.max(np.abs(S - g[f"S{l}"])) < 1e-11
        assert normerr(x, g[f"x{l}"]) < 1e-9
        uint = jitr_b200.reactions.Wavefunctions(solver, x, S, u, channels[l]).uint()[0](g["svals"])
        assert normerr(uint, g[f"uint{l}"]) < 1e-9


def test_solver_solve_coupled_config5(golden):
    """BASELINE config 5 (examples/coupled.py geometry): 3 channels x N = 40/50."""
    g = golden("config5_coupled")
    cm = g["coupling"]
    for N in (40, 50):
        solver = jitr_b200.Solver(N)
        sysc = ProjectileTargetSystem(channel_radius=float(g["a"]), lmax=3, Ztarget=20, Zproj=1,
                                      channel_levels=np.array([0, 2.3, 3.1]), coupling=lambda l: cm)
        channels, asym = sysc.get_partial_wave_channels(5.0, g["E"], g["mu"], g["k"], g["eta"])
        for l in range(4):
            R, S, x, u = solver.solve(channels[l], asym[l], local_potential=g[f"N{N}_l{l}_V"], wavefunction=True)
            assert relerr(S, g[f"N{N}_l{l}_S"]) < 1e-9 and np.max(np.abs(S - g[f"N{N}_l{l}_S"])) < RTOL_S
            assert normerr(R, g[f"N{N}_l{l}_R"]) < 1e-9
            assert normerr(x, g[f"N{N}_l{l}_x"]) < 1e-9
            assert normerr(u, g[f"N{N}_l{l}_u"]) < 1e-9
            np.testing.assert_allclose(S.conj().T @ S, np.eye(3), atol=1e-9)  # real potential: unitary


def test_solver_solve_coupled_vs_single(golden):
    """restated tests/test_coupled_single_dwba.py: zero coupling reproduces single-channel solves."""
    g = golden("test_coupled_single")
    solver = jitr_b200.Solver(40)
    sys2 = ProjectileTargetSystem(channel_radius=float(g["a"]), lmax=2, mass_target=44657, mass_projectile=938.3, Ztarget=20, Zproj=1,
                                  coupling=lambda l: np.array([[1, 0], [0, 1]]))
    channels, asym = sys2.get_partial_wave_channels(*g["kin"])
    Rm, Sm, um = solver.solve(channels[0], asym[0], local_potential=g["V2"])
    ch0, as0 = channels[0].decouple()[0], asym[0].decouple()[0]
    R1, S1, u1 = solver.solve(ch0, as0, local_potential=g["V1"])
    np.testing.assert_almost_equal(Sm[1, 0], 0)
    np.testing.assert_almost_equal(Rm[0, 1], 0)
    np.testing.assert_almost_equal(Sm[1, 1], S1[0, 0])
    np.testing.assert_almost_equal(Sm[0, 0], S1[0, 0])
    np.testing.assert_almost_equal(Rm[0, 0], R1[0, 0])
    assert np.max(np.abs(Sm - g["Sm"])) < RTOL_S
    assert relerr(S1, g["S1"]) < RTOL_S
    with pytest.raises(ValueError):
        solver.solve(channels[0], asym[0], local_potential=g["V1"])  # 1D potential with 2 channels


def test_solver_solve_nonlocal(golden):
    """nonlocal branch: Yamaguchi (examples/nonlocal.py) and the config-4 Perey-Buck-style kernel."""
    g = golden("nonlocal_yamaguchi")
    Elab, Ecm, mu, k, eta = g["kin"]
    sysn = ProjectileTargetSystem(channel_radius=30.0, lmax=5)
    channels, asym = sysn.get_partial_wave_channels(Ecm, Ecm, mu, k, eta)
    for N in (20, 40):
        solver = jitr_b200.Solver(N)
        R, S, u = solver.solve(channels[0], asym[0], nonlocal_potential=g[f"yam{N}_V"])
        assert relerr(S, g[f"yam{N}_S"]) < RTOL_S
    delta = np.rad2deg(np.real(np.log(S[0, 0]) / 2j))
    assert abs(delta - (-66.98635852)) < 1e-6
    g = golden("config4_nonlocal")
    solver = jitr_b200.Solver(60)
    sys4 = ProjectileTargetSystem(float(g["a"]), 6, Ztarget=82, Zproj=0)
    channels, asym = sys4.get_partial_wave_channels(*g["kin"])
    for l in (0, 3):
        R, S, u = solver.solve(channels[l], asym[l], nonlocal_potential=g[f"s0_l{l}_U"])
        assert relerr(S, g[f"s0_l{l}_S"]) < RTOL_S
    # batched form: the two l = 0 kernels of the two samples in one launch
    U = torch.tensor(np.stack([g["s0_l0_U"], g["s1_l0_U"]]), dtype=torch.complex128, device="cuda")
    Elab, Ecm, mu, k, eta = g["kin"]
    free = torch.tensor(solver.free_matrix(float(g["a"]), [0], [Ecm], [mu]), device="cuda")
    a0 = asym[0]
    out = solver.solve_batch(float(g["a"]), Ecm, k, 1, free, torch.tensor(np.stack([a0.Hp, a0.Hm, a0.Hpp, a0.Hmp]), device="cuda"),
                             nonlocal_potential=U)
    assert relerr(out["S"][0].cpu().numpy(), g["s0_l0_S"]) < RTOL_S
    assert relerr(out["S"][1].cpu().numpy(), g["s1_l0_S"]) < RTOL_S


def test_singular_system_reports_status():
    """An all-NaN potential must flag the system instead of crashing or hanging (numba raises LinAlgError)."""
    rx = ElasticReaction((40, 20), (1, 0), mass_target=37214.7)
    ws = jitr_b200.IntegralWorkspace(rx, rx.kinematics(10.0), 10.0, jitr_b200.Solver(20), 4)
    c = torch.full((2, 20), float("nan"), dtype=torch.complex128, device="cuda")
    c[1] = -40.0
    S, nl, status = ws.smatrix_batch(c)
    assert int(status[0]) != _cabi.STATUS_OK and int(status[1]) == _cabi.STATUS_OK
    # neighbours of a non-finite system are untouched, on the symmetric path and on partial pivoting
    lib = _cabi.lib()
    big = torch.full((64, 20), -40.0 - 2.0j, dtype=torch.complex128, device="cuda")
    big *= torch.linspace(0.8, 1.2, 64, device="cuda")[:, None]
    clean = ws.smatrix_batch(big)[0].clone()
    dirty_in = big.clone()
    dirty_in[::5] = float("nan")
    dirty_in[7, 3] = float("inf")
    keep = [i for i in range(64) if i % 5 and i != 7]
    try:
        for opt in (1, 0):
            lib.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, opt)
            ref = ws.smatrix_batch(big)[0]
            Sd, _, st = ws.smatrix_batch(dirty_in)
            assert all(int(st[i]) != _cabi.STATUS_OK for i in range(0, 64, 5)) and int(st[7]) != _cabi.STATUS_OK
            assert all(int(st[i]) == _cabi.STATUS_OK for i in keep)
            assert float((Sd[keep] - ref[keep]).abs().max()) == 0.0
    finally:
        lib.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, 1)
    assert float((clean - ws.smatrix_batch(big)[0]).abs().max()) == 0.0
    with pytest.raises(np.linalg.LinAlgError):
        ws.smatrix(np.full(20, np.nan, dtype=complex))


# ------------------------------------------------------------------------------------------------
# quasi-elastic (p,n) DWBA workspace (xs/quasielastic_pn.py)
# ------------------------------------------------------------------------------------------------
def make_qe_ws(g, tol=None):
    from jitr_b200 import reactions
    from jitr_b200.xs.quasielastic_pn import Workspace

    AZ, m = g["AZ"], g["masses"]
    rx = reactions.Reaction((int(AZ[0]), int(AZ[1])), (int(AZ[2]), int(AZ[3])), (int(AZ[4]), int(AZ[5])), (int(AZ[6]), int(AZ[7])),
                            mass_target=m[0], mass_projectile=m[1], mass_product=m[2], mass_residual=m[3])
    ke = rx.kinematics(float(g["Elab"]))
    kx = rx.kinematics_exit(ke, float(g["Ex"]))
    return Workspace(rx, ke, kx, jitr_b200.Solver(int(g["N"])), g["angles"], int(g["lmax"]), float(g["Rch"]),
                     tmatrix_abs_tol=float(g["tol"]) if tol is None else tol)


@pytest.mark.parametrize("name", ["qe_reftest", "qe_48Ca_35MeV", "qe_48Ca_35MeV_nobreak"])
def test_quasielastic_pn_vs_reference_golden(golden, name):
    """Distorted waves of both channels with coefficients, T_lj overlap, early break and the angular sum against
    the REAL reference: S to 1e-10, T_lj to 1e-9 of its largest element, dsigma/dOmega to 1e-8."""
    g = golden(name)
    ws = make_qe_ws(g)
    U = [torch.tensor(np.ascontiguousarray(g["U"][:, i]), device="cuda") for i in range(5)]
    so = bool(g["has_so"][0])
    args = (U[0], U[1], U[2] if so else None, U[3], U[4] if so else None)
    T, Sn, Sp, status = ws.tmatrix_batch(*args)
    xs = ws.xs_batch(*args)
    assert int(status.abs().sum()) == 0
    for b in range(g["U"].shape[0]):
        assert normerr(T[b].cpu().numpy(), g["Tpn"][b]) < 1e-9
        assert np.abs(Sn[b].cpu().numpy() - g["Sn"][b]).max() < RTOL_S
        assert np.abs(Sp[b].cpu().numpy() - g["Sp"][b]).max() < RTOL_S
        assert relerr(xs[b].cpu().numpy(), g["xs"][b]) < RTOL_XS
    # the break really happened where the reference's did (entries beyond it are exactly zero)
    assert np.array_equal(T.cpu().numpy() == 0, g["Tpn"] == 0)


def test_quasielastic_pn_single_sample_api_and_optional_spin_orbit(golden):
    """The reference's own test (tests/test_xs_optional_spin_orbit.py:56-103): omitted spin-orbit terms equal
    explicit zeros; numpy in, numpy out, same values as the reference."""
    g = golden("qe_reftest")
    ws = make_qe_ws(g)
    U = g["U"][0]
    zero = np.zeros_like(U[0])
    om = ws.tmatrix(U[0], U[1], U_n_central=U[3])
    ex = ws.tmatrix(U[0], U[1], zero, U[3], zero)
    for a, b in zip(om, ex):
        np.testing.assert_allclose(a, b)
    np.testing.assert_allclose(ws.xs(U[0], U[1], U_n_central=U[3]), ws.xs(U[0], U[1], zero, U[3], zero))
    assert normerr(om[0], g["Tpn"][0]) < 1e-9
    assert relerr(ws.xs(U[0], U[1], U_n_central=U[3]), g["xs"][0]) < RTOL_XS
    with pytest.raises(TypeError):
        ws.tmatrix(U[0], U[1])
    with pytest.raises(ValueError):
        ws.tmatrix(U[0][:-1], U[1], U_n_central=U[3])


def test_quasielastic_pn_vs_oracle_random_batch(golden):
    """Seeded potentials outside the golden set, mixed early breaks inside one batch, against the CPU oracle."""
    g = golden("qe_48Ca_35MeV")
    ws = make_qe_ws(g, tol=1e-5)
    AZ = g["AZ"]
    ow = orc.QuasielasticWorkspace(int(g["N"]), int(g["lmax"]), float(g["Rch"]), g["kin_entrance"], g["kin_exit"], g["angles"],
                                   int(AZ[0]), int(AZ[1]), int(AZ[1] * AZ[3]), 0, 1e-5)
    rng = np.random.default_rng(7)
    B = 6
    U = np.repeat(g["U"][:1], B, axis=0).copy()  # (B, 5, N)
    U[:, 1] *= (1 + 0.08 * rng.standard_normal((B, 1)))
    U[:, 3] *= (1 + 0.08 * rng.standard_normal((B, 1)))
    U[:, 2] *= (1 + 0.2 * rng.standard_normal((B, 1)))
    U[3:, 3] = U[3:, 1] * (1 + 1e-4)  # nearly equal p and n potentials: tiny T, early break much sooner
    Ut = [torch.tensor(np.ascontiguousarray(U[:, i]), device="cuda") for i in range(5)]
    T, Sn, Sp, status = ws.tmatrix_batch(*Ut)
    xs = ws.xs_batch(*Ut)
    assert int(status.abs().sum()) == 0
    breaks = set()
    for b in range(B):
        To, Sno, Spo = ow.tmatrix(*U[b])
        breaks.add(int(np.max(np.nonzero(np.abs(To).sum(axis=1))[0])))
        assert np.array_equal(T[b].cpu().numpy() == 0, To == 0)
        assert normerr(T[b].cpu().numpy(), To) < 1e-9
        assert np.abs(Sn[b].cpu().numpy() - Sno).max() < RTOL_S
        assert np.abs(Sp[b].cpu().numpy() - Spo).max() < RTOL_S
        assert relerr(xs[b].cpu().numpy(), ow.xs(*U[b])) < RTOL_XS
    assert len(breaks) > 1


# ------------------------------------------------------------------------------------------------
# edge cases: empty batches, a single partial wave, a single angle, the smallest and largest meshes
# ------------------------------------------------------------------------------------------------
def test_edge_cases_empty_and_minimal_shapes():
    rx = ElasticReaction((40, 20), (1, 0), mass_target=37214.7, mass_projectile=939.5654983938299)
    kin = rx.kinematics(14.0)
    model = LocalOpticalPotential()
    p = np.array([[45.0, 1.25, 0.65, 1.0, 1.25, 0.65, 6.0, 0.0, 1.25, 0.65, 5.5, 0.0, 1.1, 0.65, 1.3]])
    # empty batch: well-formed empty outputs, no launch error
    ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, 9.0, jitr_b200.Solver(24), 6, np.linspace(0.3, 2.8, 5))
    x = ws._as_sweep().xs_params(model, torch.zeros((0, 15), dtype=torch.float64, device="cuda"))
    assert tuple(x.dsdo.shape) == (0, 1, 5) and tuple(x.rxn.shape)[0] == 0
    S, nl, st = ws.integral_workspace.smatrix_batch(torch.zeros((0, 24), dtype=torch.complex128, device="cuda"))
    assert S.shape[0] == 0 and nl.shape[0] == 0
    r0 = jitr_b200.Solver(24).solve_batch(7.0, 13.6, 0.8, 1, torch.zeros((24, 24), dtype=torch.complex128, device="cuda"),
                                          torch.ones((4, 1), dtype=torch.complex128, device="cuda"),
                                          local_potential=torch.zeros((0, 24), dtype=torch.complex128, device="cuda"))
    assert r0["S"].shape[0] == 0
    # lmax = 0 (one partial wave, j = 1/2 only), one angle; smallest useful mesh (N = 3) and the largest fused one (N = 64)
    for N, lmax, nth in [(3, 0, 1), (5, 2, 1), (64, 0, 2), (24, 0, 1)]:
        ang = np.linspace(0.7, 2.1, nth)
        w = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, 9.0, jitr_b200.Solver(N), lmax, ang)
        xs = w._as_sweep().xs_params(model, torch.tensor(p, device="cuda"))
        ow = orc.ElasticWorkspace(N, lmax, 9.0, tuple(kin), ang, 0)
        c, so, co = orc.local_omp_potentials(w.radial_grid(), 40, 0, p[0])
        dsdo, Ay, Q, tot, rxn = ow.xs(c, so, co)
        assert relerr(xs.dsdo[0, 0].cpu().numpy(), dsdo) < RTOL_XS, (N, lmax)
        assert abs(float(xs.t[0, 0]) - tot) <= RTOL_XS * abs(tot), (N, lmax)


# ------------------------------------------------------------------------------------------------
# blocked CTA-per-system kernel: symmetric (diagonal-pivot) path, its fall-back, asymmetric couplings
# ------------------------------------------------------------------------------------------------
def _coupled_case(rng, B, N, nch, W, symmetric=True):
    solver = jitr_b200.Solver(N)
    a, E0, k0 = 6 * np.pi, 4.9, 0.48
    E = [4.9, 2.6, 1.8][:nch]
    l = [0, 2, 2][:nch]
    free = solver.free_matrix(a, l, E, [919.0] * nch)
    r = solver.radial_grid(a, k0)
    V = np.zeros((B, nch, nch, N), dtype=np.complex128)
    ws = -(42.0 + 1j * W) / (1 + np.exp((r - 4.0) / 0.8))
    for i in range(nch):
        V[:, i, i] = ws[None] * (1 + 0.05 * rng.standard_normal((B, 1)))
        for j in range(i + 1, nch):
            c = 5.0 * np.exp(-((r - 4.0) / 1.2) ** 2)
            V[:, i, j] = c[None] * (1 + 0.05 * rng.standard_normal((B, 1)))
            V[:, j, i] = V[:, i, j] if symmetric else 0.7 * V[:, i, j]
    asym = np.array([orc.coulomb_hankel(li, 0.0, a) for li in l]).T  # (4, nch)
    return solver, a, E0, k0, l, E, free, V, asym


def _oracle_coupled(N, a, l, E, V, asym, k0):
    """A = F + V / E0 and the reference's solve (rmatrix/core.py) for one coupled set."""
    nch = len(l)
    F = orc.free_matrix(N, a, l, E, [919.0] * nch)
    Vm = orc.interaction_matrix(N, k0, E[0], a, nch, local_potential=V)
    w = np.zeros(nch)
    w[0] = 1
    R, S, _, _ = orc.solve_smatrix(F + Vm, orc.boundary_vector(N), *[np.ascontiguousarray(asym[i]) for i in range(4)], w, a, nch, N)
    return R, S


@pytest.mark.parametrize("N,nch", [(20, 2), (21, 2), (37, 3), (50, 3)])
def test_blocked_solver_symmetric_path_and_fallback(N, nch):
    lib = _cabi.lib()
    rng = np.random.default_rng(11)
    B = 64
    try:
        for W, expect_fallbacks in [(4.0, False), (0.0, None)]:
            solver, a, E0, k0, l, E, free, V, asym = _coupled_case(rng, B, N, nch, W)
            Ft, Vt, At = torch.tensor(free, device="cuda"), torch.tensor(V, device="cuda"), torch.tensor(asym, device="cuda")
            lib.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, 1)
            # a first pass with other data leaves stale numbers in the slabs (padding columns must not leak in)
            solver.solve_batch(a, E0, k0, nch, Ft * 1.3, At, local_potential=Vt * (0.7 + 0.2j))
            f0 = lib.jitr_b200_sym_fallback_count()
            rs = solver.solve_batch(a, E0, k0, nch, Ft, At, local_potential=Vt)
            torch.cuda.synchronize()
            nfb = lib.jitr_b200_sym_fallback_count() - f0
            rsw = solver.solve_batch(a, E0, k0, nch, Ft, At, local_potential=Vt, wavefunction=True)
            torch.cuda.synchronize()
            f1 = lib.jitr_b200_sym_fallback_count()
            lib.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, 0)
            rp = solver.solve_batch(a, E0, k0, nch, Ft, At, local_potential=Vt)
            rpw = solver.solve_batch(a, E0, k0, nch, Ft, At, local_potential=Vt, wavefunction=True)
            torch.cuda.synchronize()
            assert lib.jitr_b200_sym_fallback_count() == f1  # partial pivoting only: no attempts
            # wavefunction coefficients from the LDL^T factors (U = D L^T stored transposed) = those from the LU
            assert float((rsw["S"] - rp["S"]).abs().max()) < RTOL_S
            assert normerr(rsw["coeffs"].cpu().numpy(), rpw["coeffs"].cpu().numpy()) < 1e-9
            assert int(rs["status"].abs().sum()) == 0 and int(rp["status"].abs().sum()) == 0
            assert float((rs["S"] - rp["S"]).abs().max()) < RTOL_S
            if expect_fallbacks is False:
                assert nfb == 0  # an absorptive potential keeps the diagonal pivots away from zero
            for b in (0, B - 1):
                R, S = _oracle_coupled(N, a, l, E, V[b], asym, k0)
                assert np.abs(rs["S"][b].cpu().numpy() - S).max() < RTOL_S
                assert normerr(rs["R"][b].cpu().numpy(), R) < 1e-9
    finally:
        lib.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, 1)


def test_blocked_solver_asymmetric_couplings_take_partial_pivoting():
    """V_ij != V_ji: the in-kernel symmetry check must route the system to the partial-pivoting path."""
    rng = np.random.default_rng(12)
    N, nch, B = 20, 2, 8
    solver, a, E0, k0, l, E, free, V, asym = _coupled_case(rng, B, N, nch, 3.0, symmetric=False)
    r = solver.solve_batch(a, E0, k0, nch, torch.tensor(free, device="cuda"), torch.tensor(asym, device="cuda"),
                           local_potential=torch.tensor(V, device="cuda"))
    assert int(r["status"].abs().sum()) == 0
    for b in range(B):
        R, S = _oracle_coupled(N, a, l, E, V[b], asym, k0)
        assert np.abs(r["S"][b].cpu().numpy() - S).max() < RTOL_S


# ------------------------------------------------------------------------------------------------
# dispersive optical model: host parameter map per (sample, energy) + device forms (MODEL_SHAPE_PAIR)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("host_map", [False, True])
@pytest.mark.parametrize("name", ["dom_n40Ca_14MeV", "dom_p208Pb_30MeV"])
def test_dom_sweep_vs_reference_golden(golden, name, host_map):
    """host_map = False: the parameter map (incl. the exponential integrals of the surface-dispersion term) runs inside the fused
    kernel from the 19 global parameters; True: the round-1 route, the map on the host, one canonical row per (sample, energy)."""
    from jitr_b200.optical_potentials import dom

    g = golden(name)
    rx = ElasticReaction(tuple(int(v) for v in g["target"]), tuple(int(v) for v in g["projectile"]),
                         mass_target=float(g["masses"][0]), mass_projectile=float(g["masses"][1]), Ef=float(g["Ef"]))
    kin = rx.kinematics(float(g["kin"][0]))
    ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, float(g["Rch"]), jitr_b200.Solver(int(g["N"])), int(g["lmax"]), g["angles"])
    model = dom.DOM(host_map=host_map)
    x, status = ws._as_sweep().xs_params(model, g["params"], return_status=True)
    S, nl, _ = ws.integral_workspace.smatrix_params(model, g["params"])
    assert int(status.abs().sum()) == 0
    for i in range(g["params"].shape[0]):
        L = int(g["nl"][i])
        assert int(nl[i]) == L
        assert relerr(S[i, 0, :L].cpu().numpy(), g["Splus"][i, :L]) < RTOL_S
        assert relerr(S[i, 1, 1:L].cpu().numpy(), g["Sminus"][i, 1:L]) < RTOL_S
        assert relerr(x.dsdo[i, 0].cpu().numpy(), g["dsdo"][i]) < RTOL_XS
        assert abs(float(x.rxn[i, 0]) - g["rxn"][i]) <= RTOL_XS * abs(g["rxn"][i])
    # the reference's single-sample contract: evaluate(rgrid, reaction, kinematics, *params) -> mesh arrays
    c, so, co = model(g["rgrid"], rx, kin, *g["params"][2])
    assert normerr(c, g["central"][2]) < 1e-12 and normerr(so, g["so"][2]) < 1e-12
    assert np.max(np.abs(co - g["coul"][2])) <= 1e-12 * max(1.0, np.max(np.abs(g["coul"][2])))
    with pytest.raises(ValueError):
        model(g["rgrid"], rx, kin, *g["params"][2][:-1])


@pytest.mark.parametrize("host_map", [False, True])
def test_dom_multi_energy_sweep_vs_oracle(host_map):
    """Rows differ per energy: a three-energy sweep of the same samples against the oracle energy by energy."""
    from jitr_b200.optical_potentials import dom

    g = np.load("tests/golden/dom_n40Ca_14MeV.npz")
    rx = ElasticReaction((40, 20), (1, 0), mass_target=float(g["masses"][0]), mass_projectile=float(g["masses"][1]), Ef=float(g["Ef"]))
    ang = np.linspace(0.2, 3.0, 31)
    solver = jitr_b200.Solver(32)
    wss = [jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(E), 10.5, solver, 16, ang) for E in (6.0, 14.0, 27.0)]
    sweep = jitr_b200.ElasticSweep(wss)
    x = sweep.xs_params(dom.DOM(host_map=host_map), g["params"][:3])
    for e, w in enumerate(wss):
        ow = orc.ElasticWorkspace(32, 16, 10.5, tuple(w.kinematics), ang, 0)
        for b in range(3):
            c, so, co = orc.dom_potentials(w.radial_grid(), (1, 0), (40, 20), w.kinematics.Ecm, float(g["Ef"]), g["params"][b])
            dsdo, Ay, Q, tot, rxn = ow.xs(c, so, co)
            assert relerr(x.dsdo[b, e].cpu().numpy(), dsdo) < RTOL_XS
            assert abs(float(x.rxn[b, e]) - rxn) <= RTOL_XS * abs(rxn)


def test_dom_mixed_target_sweep_equals_per_workspace_results():
    """A DOM sweep whose energy rows belong to DIFFERENT targets (and Fermi energies): every row must use its own A, Z, Ef,
    Coulomb correction and radii -- the sweep equals the single-workspace results row by row, and the oracle."""
    from jitr_b200.optical_potentials import dom

    g = np.load("tests/golden/dom_n40Ca_14MeV.npz")
    ang = np.linspace(0.2, 3.0, 31)
    solver = jitr_b200.Solver(32)
    rx_ca = ElasticReaction((40, 20), (1, 0), mass_target=float(g["masses"][0]), mass_projectile=float(g["masses"][1]), Ef=float(g["Ef"]))
    rx_pb = ElasticReaction((208, 82), (1, 0), mass_target=193687.12278341982, mass_projectile=float(g["masses"][1]), Ef=-5.65)
    wss = [jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(E), R, solver, 16, ang)
           for rx, E, R in ((rx_ca, 9.0, 10.5), (rx_pb, 9.0, 13.0), (rx_ca, 21.0, 10.5), (rx_pb, 21.0, 13.0))]
    p = g["params"][:3]
    xh = jitr_b200.ElasticSweep(wss).xs_params(dom.DOM(host_map=True), p)
    model = dom.DOM()
    x = jitr_b200.ElasticSweep(wss).xs_params(model, p)
    assert relerr(x.dsdo.cpu().numpy(), xh.dsdo.cpu().numpy()) < 1e-9   # device map vs host map
    for e, w in enumerate(wss):
        one = jitr_b200.ElasticSweep([w]).xs_params(model, p)
        assert torch.equal(one.dsdo[:, 0], x.dsdo[:, e]) and torch.equal(one.rxn[:, 0], x.rxn[:, e])
        tgt = (w.reaction.target.A, w.reaction.target.Z)
        ow = orc.ElasticWorkspace(32, 16, 10.5 if tgt[0] == 40 else 13.0, tuple(w.kinematics), ang, 0)
        c, so, co = orc.dom_potentials(w.radial_grid(), (1, 0), tgt, w.kinematics.Ecm, float(w.reaction.Ef), p[1])
        assert relerr(x.dsdo[1, e].cpu().numpy(), ow.xs(c, so, co)[0]) < RTOL_XS


def test_blocked_solver_many_channels_two_border_tiles():
    """nch = 9 > 8: two tiles of border rows / right-hand sides, on the symmetric path, on partial pivoting, and with
    coefficients, against the reference algorithm."""
    lib = _cabi.lib()
    rng = np.random.default_rng(21)
    N, nch, B = 6, 9, 5
    solver = jitr_b200.Solver(N)
    a, k0 = 9.0, 0.7
    E = list(12.0 - 0.8 * np.arange(nch))
    l = [int(v) for v in (np.arange(nch) % 3)]
    mu = [900.0] * nch
    free = solver.free_matrix(a, l, E, mu)
    V = np.zeros((B, nch, nch, N), dtype=np.complex128)
    for i in range(nch):
        for j in range(i, nch):
            v = (rng.standard_normal((B, N)) + 0.4j * rng.standard_normal((B, N))) * (20.0 if i == j else 3.0) - (25.0 + 3j) * (i == j)
            V[:, i, j] = v
            V[:, j, i] = v
    asym = np.array([orc.coulomb_hankel(li, 0.0, a) for li in l]).T
    Ft, Vt, At = torch.tensor(free, device="cuda"), torch.tensor(V, device="cuda"), torch.tensor(asym, device="cuda")
    try:
        for opt in (1, 0):
            lib.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, opt)
            r = solver.solve_batch(a, E[0], k0, nch, Ft, At, local_potential=Vt, wavefunction=True)
            assert int(r["status"].abs().sum()) == 0
            for b in range(B):
                Fm = orc.free_matrix(N, a, l, E, mu)
                Vm = orc.interaction_matrix(N, k0, E[0], a, nch, local_potential=V[b])
                w = np.zeros(nch)
                w[0] = 1
                R, S, Ainv, up = orc.solve_smatrix(Fm + Vm, orc.boundary_vector(N), *[np.ascontiguousarray(asym[i]) for i in range(4)], w, a, nch, N)
                x = orc.solution_coeffs(Ainv, orc.boundary_vector(N), up, nch, N)
                assert np.abs(r["S"][b].cpu().numpy() - S).max() < RTOL_S
                assert normerr(r["R"][b].cpu().numpy(), R) < 1e-9
                assert normerr(r["coeffs"][b].cpu().numpy(), x) < 1e-9
    finally:
        lib.jitr_b200_set_option(_cabi.OPT_SYMMETRIC, 1)


@pytest.mark.parametrize("N,nch,kind", [(3, 2, "local"), (5, 16, "local"), (60, 4, "local"), (7, 3, "nonlocal_sym"), (7, 3, "nonlocal_asym"),
                                        (9, 2, "interaction"), (6, 3, "per_system_free")])
def test_blocked_solver_shapes_and_inputs_vs_oracle(N, nch, kind):
    """Corners of the blocked kernel: fewer rows than one panel, 16 channels, 248 bordered rows, coupled nonlocal
    kernels (symmetric and not), a ready interaction matrix, one free matrix per system -- all with coefficients."""
    rng = np.random.default_rng(100 + N + nch)
    B = 3
    solver = jitr_b200.Solver(N)
    a, k0 = 8.0, 0.9
    E = list(10.0 - 0.45 * np.arange(nch))
    l = [int(v) for v in (np.arange(nch) % 3)]
    mu = [930.0] * nch
    n = N * nch
    free = solver.free_matrix(a, l, E, mu)
    asym = np.array([orc.coulomb_hankel(li, 0.0, a) for li in l]).T
    cplx = lambda *sh: rng.standard_normal(sh) + 0.3j * rng.standard_normal(sh)  # noqa: E731
    loc = np.zeros((B, nch, nch, N), dtype=np.complex128)
    for i in range(nch):
        for j in range(i, nch):
            v = cplx(B, N) * (8.0 if i == j else 1.5) - (20.0 + 2j) * (i == j)
            loc[:, i, j] = v
            loc[:, j, i] = v
    kw, nl, Fb = {}, None, None
    if kind.startswith("nonlocal"):
        nl = cplx(B, nch, nch, N, N) * 3.0
        if kind == "nonlocal_sym":
            nl = 0.5 * (nl + nl.transpose(0, 2, 1, 4, 3))
        kw = dict(local_potential=torch.tensor(loc, device="cuda"), nonlocal_potential=torch.tensor(nl, device="cuda"))
    elif kind == "interaction":
        kw = dict(interaction_matrix=torch.tensor(np.stack([orc.interaction_matrix(N, k0, E[0], a, nch, local_potential=loc[b]) for b in range(B)]), device="cuda"))
    else:
        kw = dict(local_potential=torch.tensor(loc, device="cuda"))
    if kind == "per_system_free":
        Fb = np.stack([free * (1 + 0.01 * b) for b in range(B)])
    Ft = torch.tensor(free if Fb is None else Fb, device="cuda")
    r = solver.solve_batch(a, E[0], k0, nch, Ft, torch.tensor(asym, device="cuda"), wavefunction=True, batch=None if kind != "interaction" else B, **kw)
    assert int(r["status"].abs().sum()) == 0
    w = np.zeros(nch)
    w[0] = 1
    bvec = orc.boundary_vector(N)
    for b in range(B):
        Fm = orc.free_matrix(N, a, l, E, mu) if Fb is None else Fb[b]
        Vm = orc.interaction_matrix(N, k0, E[0], a, nch, local_potential=loc[b], nonlocal_potential=None if nl is None else nl[b])
        R, S, Ainv, up = orc.solve_smatrix(Fm + Vm, bvec, *[np.ascontiguousarray(asym[i]) for i in range(4)], w, a, nch, N)
        x = orc.solution_coeffs(Ainv, bvec, up, nch, N)
        assert np.abs(r["S"][b].cpu().numpy() - S).max() < RTOL_S, (kind, b)
        assert normerr(r["R"][b].cpu().numpy(), R) < 1e-9
        assert normerr(r["coeffs"][b].cpu().numpy(), x) < 1e-9
        assert r["S"].shape == (B, nch, nch) and r["coeffs"].shape == (B, nch, N)
    assert n == N * nch


@pytest.mark.parametrize("N,nch", [(30, 1), (56, 1), (50, 3)])
def test_general_solver_status_codes(N, nch):
    """Non-finite and singular systems are flagged per system (numba's LinAlgError in the reference), on the
    warp-per-system kernel (N = 30) and on both CTA sizes of the blocked kernel; their neighbours are unaffected."""
    rng = np.random.default_rng(5)
    solver = jitr_b200.Solver(N)
    a, k0 = 9.0, 0.8
    E = [9.0, 7.5, 6.0][:nch]
    l = [0, 1, 2][:nch]
    free = solver.free_matrix(a, l, E, [930.0] * nch)
    asym = np.array([orc.coulomb_hankel(li, 0.0, a) for li in l]).T
    B = 6
    V = np.zeros((B, nch, nch, N), dtype=np.complex128)
    for i in range(nch):
        V[:, i, i] = -(30.0 + 3j) * (1 + 0.1 * rng.standard_normal((B, 1))) / (1 + np.exp((np.linspace(0.3, 11, N) - 4.5) / 0.6))
    V[1] = np.nan
    V[4, 0, 0, N // 2] = np.inf
    Ft = torch.tensor(free, device="cuda")
    r = solver.solve_batch(a, E[0], k0, nch, Ft, torch.tensor(asym, device="cuda"), local_potential=torch.tensor(V, device="cuda"))
    st = r["status"].cpu().numpy()
    assert st[1] != _cabi.STATUS_OK and st[4] != _cabi.STATUS_OK
    assert all(st[b] == _cabi.STATUS_OK for b in (0, 2, 3, 5))
    good = solver.solve_batch(a, E[0], k0, nch, Ft, torch.tensor(asym, device="cuda"), local_potential=torch.tensor(V[[0, 2, 3, 5]], device="cuda"))
    assert float((r["S"][[0, 2, 3, 5]] - good["S"]).abs().max()) == 0.0
    # exactly singular: a zero matrix (free matrix cancelled by the interaction)
    zero = solver.solve_batch(a, E[0], k0, nch, Ft, torch.tensor(asym, device="cuda"), interaction_matrix=(-Ft)[None].contiguous(), batch=1)
    assert int(zero["status"][0]) != _cabi.STATUS_OK


def test_two_devices_from_two_threads_in_one_process(golden):
    """The library keeps its state per device (SM count, work counters, fall-back counter, kernel attributes): two
    threads drive cuda:0 and cuda:1 at once and both reproduce the reference.  Skipped on a single-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import threading

    g0 = golden("config1_n208Pb_30MeV")
    g = {k: g0[k] for k in g0.files}  # NpzFile reads lazily and is not thread-safe
    out, err = {}, []

    def work(dev):
        try:
            with torch.cuda.device(dev):
                rx, kin, ws = make_ws(g)
                model = model_for("config1", g)
                for _ in range(3):
                    x = ws._as_sweep().xs_params(model, torch.tensor(np.tile(g["params"][:1], (4096, 1)), device=f"cuda:{dev}"))
                N = int(g["N"])
                solver = jitr_b200.Solver(N)
                free = torch.tensor(solver.free_matrix(12.0, [0]), device=f"cuda:{dev}")
                asym = torch.tensor(np.array(orc.coulomb_hankel(0, 0.0, 12.0)).reshape(4, 1), device=f"cuda:{dev}")
                V = torch.full((256, 1, 1, N), -30.0 - 2j, dtype=torch.complex128, device=f"cuda:{dev}")
                r = solver.solve_batch(12.0, 20.0, 1.0, 1, free, asym, local_potential=V, wavefunction=True)
                torch.cuda.synchronize(dev)
                out[dev] = (x.dsdo[0, 0].cpu().numpy(), r["S"].cpu().numpy())
        except Exception as e:  # noqa: BLE001
            err.append((dev, repr(e)))

    ts = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not err, err
    assert relerr(out[0][0], g["dsdo"][0]) < RTOL_XS and relerr(out[1][0], g["dsdo"][0]) < RTOL_XS
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_large_lmax_proton_and_odd_batch_sizes():
    """lmax = 100 (the reference's Frescox regression cases go that high) for a charged projectile -- Coulomb-Hankel
    table up to l = 101, early break far below lmax -- and batch sizes that do not fill the last CTA / warp."""
    rx = ElasticReaction((58, 28), (1, 1), mass_target=53966.4, mass_projectile=938.271653086152)
    kin = rx.kinematics(35.0)
    ang = np.linspace(0.15, 3.0, 11)
    N, lmax = 48, 100
    ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, 14.0, jitr_b200.Solver(N), lmax, ang)
    model = LocalOpticalPotential()
    base = np.array([50.0, 1.2, 0.7, 3.0, 1.25, 0.6, 6.5, 0.0, 1.25, 0.6, 5.9, -0.2, 1.05, 0.65, 1.25])
    ow = orc.ElasticWorkspace(N, lmax, 14.0, tuple(kin), ang, 28)
    r = ws.radial_grid()
    for B in (1, 13, 385):
        rng = np.random.default_rng(B)
        p = np.tile(base, (B, 1)) * (1 + 0.03 * rng.standard_normal((B, base.size)))
        x, status = ws._as_sweep().xs_params(model, torch.tensor(p, device="cuda"), return_status=True)
        S, nl, _ = ws.integral_workspace.smatrix_params(model, torch.tensor(p, device="cuda"))
        assert int(status.abs().sum()) == 0
        for b in sorted({0, B // 2, B - 1}):
            c, so, co = orc.local_omp_potentials(r, 58, 28, p[b])
            sp, sm = ow.smatrix(c, so, co)
            assert int(nl[b]) == len(sp) and len(sp) < lmax
            assert relerr(S[b, 0, : len(sp)].cpu().numpy(), sp) < RTOL_S
            dsdo = ow.xs(c, so, co)[0]
            assert relerr(x.dsdo[b, 0].cpu().numpy(), dsdo) < RTOL_XS


def test_device_misfit_matches_notebook_log_likelihood(golden):
    """(f)4: the batched likelihood.  The reference's quickstart notebook (cells 11-13) forms
    y_pred = dsdo / rutherford, chi2 = sum((y_pred - y)**2 / y_err * 2), logl = -0.5 (chi2 + sum(log(2 pi y_err)));
    here chi2 is reduced inside the observable kernel (weights 2 / y_err, scale 1 / rutherford) without writing the
    angular distributions.  Checked against the oracle's dsdo for a proton case, and against the kernel's own dsdo."""
    g = golden("config3_p48Ca_30MeV")
    rx, kin, ws = make_ws(g)
    model = model_for("config3", g)
    sweep = ws._as_sweep()
    B = 24
    rng = np.random.default_rng(4)
    p = np.tile(g["params"][:1], (B, 1)) * (1 + 0.02 * rng.standard_normal((B, g["params"].shape[1])))
    p[:, -2:] = g["params"][0, -2:]
    ruth = np.asarray(ws.rutherford, dtype=np.float64)
    y = (g["dsdo"][0] / ruth) * (1 + 0.05 * rng.standard_normal(ruth.shape))
    y_err = 0.05 * np.abs(y) + 1e-3
    mis, tot, rxn, status = sweep.misfit_params(model, torch.tensor(p, device="cuda"), y[None], (2.0 / y_err)[None], (1.0 / ruth)[None])
    x = sweep.xs_params(model, torch.tensor(p, device="cuda"))
    assert int(status.abs().sum()) == 0
    dsdo = x.dsdo[:, 0].cpu().numpy()
    chi2_k = np.sum((dsdo / ruth - y) ** 2 / y_err * 2, axis=1)
    assert relerr(mis[:, 0].cpu().numpy(), chi2_k) < 1e-12
    assert float((rxn - x.rxn).abs().max()) == 0.0
    ow = orc.ElasticWorkspace(int(g["N"]), int(g["lmax"]), float(g["Rch"]), tuple(g["kin"]), g["angles"], 20, asym_table=g["asym"])
    log_err_term = np.sum(np.log(2 * np.pi * y_err))
    for b in (0, B - 1):
        c, so, co = orc.kduq_potentials(ws.radial_grid(), (1, 1), (48, 20), float(g["kin"][0]), p[b])
        y_pred = ow.xs(c, so, co)[0] / ruth
        chi2 = np.sum((y_pred - y) ** 2 / y_err * 2)
        logl_ref = -0.5 * (chi2 + log_err_term)
        logl = -0.5 * (float(mis[b, 0]) + log_err_term)
        assert abs(logl - logl_ref) <= 1e-8 * abs(logl_ref)
    # with the angular distributions requested as well: identical numbers to the observable kernel's
    S, nl, _ = sweep.smatrix_params(model, torch.tensor(p, device="cuda"))
    m2, _, _, d2 = sweep.misfit(S, nl, y[None], (2.0 / y_err)[None], (1.0 / ruth)[None], want_dsdo=True)
    assert float((d2 - x.dsdo).abs().max()) == 0.0 and float((m2 - mis).abs().max()) == 0.0
    with pytest.raises(ValueError):
        sweep.misfit(S, nl, y, (2.0 / y_err)[None], None)


def test_device_percentiles_match_numpy_bit_for_bit():
    """(f)4: uncertainty bands on the device = numpy.percentile(values, q, axis=0) (kduq_uq_demo cell 18), exact order
    statistics and numpy's linear interpolation: heavy ties, negative numbers and zeros of both signs, one and two samples,
    column counts that do not fill the last CTA, q at both ends."""
    from jitr_b200 import uq

    rng = np.random.default_rng(8)
    cases = [rng.standard_normal((1000, 37)) * 10.0, np.round(rng.standard_normal((513, 16)) * 2.0) / 2.0, rng.random((1, 5)),
             rng.random((2, 33)) - 0.5, np.abs(rng.standard_normal((4097, 3))) * 1e-300, rng.standard_normal((64, 2, 7)) * 1e6]
    cases[1][::7, 3] = -0.0
    cases[1][1::7, 3] = 0.0
    for v in cases:
        for q in ([16.0, 84.0], [0.0, 2.5, 50.0, 97.5, 100.0], [33.3]):
            got = uq.percentiles(torch.tensor(v, device="cuda"), q).cpu().numpy()
            want = np.percentile(v, q, axis=0)
            assert got.shape == want.shape
            assert np.array_equal(got, want), (v.shape, q, np.abs(got - want).max())
    with pytest.raises(ValueError):
        uq.percentiles(torch.zeros((4, 4), device="cuda", dtype=torch.float64), [101.0])


def test_device_percentiles_nan_column_like_numpy():
    """A column that holds a NaN (a sample whose system was flagged non-finite) gives NaN, as numpy.percentile does; its
    neighbours in the same 16-column tile are untouched."""
    from jitr_b200 import uq

    rng = np.random.default_rng(3)
    x = rng.standard_normal((257, 40))
    x[17, 5] = np.nan
    x[200, 33] = np.nan
    x[3, 34] = -0.0
    q = [2.5, 50.0, 97.5]
    with np.errstate(invalid="ignore"):
        ref = np.percentile(x, q, axis=0)
    got = uq.percentiles(torch.tensor(x, device="cuda"), q).cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.isnan(got[:, 5]).all() and np.isnan(got[:, 33]).all()
    ok = ~np.isnan(ref)
    assert np.array_equal(got[ok], ref[ok])


def test_device_percentile_bands_of_a_sweep(golden):
    """The notebook's band computation on real sweep output: np.percentile(dsdo / rutherford, [16, 84], axis=samples)."""
    from jitr_b200 import uq

    g = golden("config3_p48Ca_30MeV")
    rx, kin, ws = make_ws(g)
    model = model_for("config3", g)
    rng = np.random.default_rng(9)
    B = 600
    p = np.tile(g["params"][:1], (B, 1)) * (1 + 0.02 * rng.standard_normal((B, g["params"].shape[1])))
    p[:, -2:] = g["params"][0, -2:]
    x = ws._as_sweep().xs_params(model, torch.tensor(p, device="cuda"))
    ratio = x.dsdo[:, 0] / torch.tensor(np.asarray(ws.rutherford), device="cuda")
    bands = uq.percentiles(ratio, [16, 84]).cpu().numpy()
    assert np.array_equal(bands, np.percentile(ratio.cpu().numpy(), [16, 84], axis=0))
    assert np.all(bands[0] <= bands[1])


# ------------------------------------------------------------------ packed symmetric single-channel solve (config 4 path)
def _single_channel_reference(N, a, l, E0, k0, asym, loc=None, nl=None, inter=None):
    F = orc.free_matrix(N, a, [l])
    V = inter if inter is not None else orc.interaction_matrix(N, k0, E0, a, 1, local_potential=loc, nonlocal_potential=nl)
    R, S, _, up = orc.solve_smatrix(F + V, orc.boundary_vector(N), *[np.ascontiguousarray(asym[i]) for i in range(4)], np.ones(1), a, 1, N)
    return R, S, up


@pytest.mark.parametrize("N", [5, 16, 23, 40, 47, 56, 60, 64])
def test_packed_single_channel_solve_vs_oracle(N):
    """jitr_b200_solve_batched, nch = 1, no coefficients: the packed LDL^T kernel with bulk-async staging of the nonlocal
    kernels (rmatrix.py:160-178, kernel.py:197-214).  Symmetric kernels stay on it, a kernel that is not symmetric in
    (r, r') is detected on the device and solved by the partial-pivoting kernel, a system whose first diagonal entry is tiny
    fails the threshold test and is deferred too; all against the oracle's explicit-inverse solve."""
    rng = np.random.default_rng(1000 + N)
    B = 37
    solver = jitr_b200.Solver(N)
    a, E0, k0, l = 11.0, 21.0, 0.95, 2
    asym = np.array([orc.coulomb_hankel(l, 0.0, a)]).T  # (4, 1)
    r = solver.radial_grid(a, k0)
    rg, rpg = np.meshgrid(r, r, indexing="ij")
    base = -38.0 * np.exp(-((rg - rpg) / 0.9) ** 2) / (1 + np.exp((0.5 * (rg + rpg) - 5.0) / 0.65)) * (1 + 0.25j)
    nl = base[None] * (1.0 + 0.05 * rng.standard_normal((B, 1, 1)))
    loc = (-4.0 - 1.0j) / (1 + np.exp((r[None] - 5.0) / 0.6)) * (1.0 + 0.1 * rng.standard_normal((B, 1)))
    asym_rows = [3, 17]      # not symmetric: must take the partial-pivoting kernel
    for b in asym_rows:
        nl[b, N // 2, 0] *= 1.0 + 1e-9
    tiny = [5]               # a^2 A[0][0] ~ 0: fails the threshold test
    Fm = orc.free_matrix(N, a, [l])
    w = solver.kernel.quadrature.weights
    for b in tiny:
        loc[b, 0] = -(Fm[0, 0] + nl[b, 0, 0] * w[0] * (a / k0) / k0 / E0) * E0 + 1e-9
    L = _cabi.lib()
    free = torch.tensor(Fm, device="cuda")
    f0 = L.jitr_b200_sym_fallback_count()
    out = solver.solve_batch(a, E0, k0, 1, free, torch.tensor(asym, device="cuda"), local_potential=torch.tensor(loc, device="cuda"),
                             nonlocal_potential=torch.tensor(nl, device="cuda"))
    assert int(out["status"].abs().sum()) == 0
    assert L.jitr_b200_sym_fallback_count() - f0 >= len(tiny)
    for b in range(B):
        R, S, up = _single_channel_reference(N, a, l, E0, k0, asym, loc=loc[b], nl=nl[b])
        assert relerr(out["S"][b].cpu().numpy(), S) < (RTOL_S if b not in tiny else 1e-7), (N, b)
        assert relerr(out["R"][b].cpu().numpy(), R) < (1e-9 if b not in tiny else 1e-6), (N, b)
        assert relerr(out["uext"][b].cpu().numpy(), up) < (1e-9 if b not in tiny else 1e-6), (N, b)
    # the caller vouches for symmetry: only the lower triangles are read, so the two doctored kernels now give the answer of
    # their lower triangle mirrored
    try:
        assert L.jitr_b200_set_option(_cabi.OPT_ASSUME_SYMMETRIC_KERNEL, 1) == 0
        out2 = solver.solve_batch(a, E0, k0, 1, free, torch.tensor(asym, device="cuda"), local_potential=torch.tensor(loc, device="cuda"),
                                  nonlocal_potential=torch.tensor(nl, device="cuda"))
    finally:
        L.jitr_b200_set_option(_cabi.OPT_ASSUME_SYMMETRIC_KERNEL, 0)
    for b in range(B):
        if b in asym_rows:
            low = np.tril(nl[b]) + np.tril(nl[b], -1).T
            R, S, up = _single_channel_reference(N, a, l, E0, k0, asym, loc=loc[b], nl=low)
            assert relerr(out2["S"][b].cpu().numpy(), S) < RTOL_S
        elif b not in tiny:
            assert torch.equal(out2["S"][b], out["S"][b])
    # wavefunction coefficients from the same factors (back substitution on the packed triangle, core.py:63-82)
    outw = solver.solve_batch(a, E0, k0, 1, free, torch.tensor(asym, device="cuda"), local_potential=torch.tensor(loc, device="cuda"),
                              nonlocal_potential=torch.tensor(nl, device="cuda"), wavefunction=True)
    assert int(outw["status"].abs().sum()) == 0
    bvec = orc.boundary_vector(N)
    for b in range(0, B, 3):
        F_ = orc.free_matrix(N, a, [l])
        V_ = orc.interaction_matrix(N, k0, E0, a, 1, local_potential=loc[b], nonlocal_potential=nl[b])
        R, S, Ainv, up = orc.solve_smatrix(F_ + V_, bvec, *[np.ascontiguousarray(asym[i]) for i in range(4)], np.ones(1), a, 1, N)
        x = orc.solution_coeffs(Ainv, bvec, up, 1, N)
        assert normerr(outw["coeffs"][b].cpu().numpy(), x) < (1e-9 if b not in tiny else 1e-6), (N, b)
        assert relerr(outw["S"][b].cpu().numpy(), S) < (RTOL_S if b not in tiny else 1e-7)
    # a ready interaction matrix, and local-only input, take the same kernel
    inter = np.stack([orc.interaction_matrix(N, k0, E0, a, 1, local_potential=loc[b], nonlocal_potential=nl[b]) for b in range(4)])
    out3 = solver.solve_batch(a, E0, k0, 1, free, torch.tensor(asym, device="cuda"), interaction_matrix=torch.tensor(inter, device="cuda"))
    out4 = solver.solve_batch(a, E0, k0, 1, free, torch.tensor(asym, device="cuda"), local_potential=torch.tensor(loc[:4], device="cuda"))
    for b in range(4):
        if b in asym_rows:
            continue
        assert relerr(out3["S"][b].cpu().numpy(), out["S"][b].cpu().numpy()) < 1e-11
        R, S, up = _single_channel_reference(N, a, l, E0, k0, asym, loc=loc[b])
        assert relerr(out4["S"][b].cpu().numpy(), S) < RTOL_S


def test_device_perey_buck_generator_vs_scipy(golden):
    """csrc/nonlocal_gen.cu against scipy's exponentially scaled Bessel function (workloads.perey_buck_kernel, the
    definition the config-4 goldens were made with), against the golden kernels of the real reference run, and against the
    reference's own perey_buck_nonlocal where that does not overflow."""
    from scipy import special as sc

    from jitr_b200 import workloads
    from jitr_b200.optical_potentials import nonlocal_forms

    g = golden("config4_nonlocal")
    solver = jitr_b200.Solver(60)
    k = float(g["kin"][3])
    rg, rpg = solver.nonlocal_radial_grids(float(g["a"]), k)
    r = solver.radial_grid(float(g["a"]), k)
    smp = workloads.config4_samples(5)
    smp[0] = (71.0, 15.0, 0.85)
    lmax = 30
    U = nonlocal_forms.perey_buck_kernels(r, lmax, smp[:, 0], smp[:, 1], smp[:, 2], R=workloads.CONFIG4["R"], a=workloads.CONFIG4["a"]).cpu().numpy()
    for b in range(5):
        for l in (0, 1, 7, 19, 30):
            ref = workloads.perey_buck_kernel(rg, rpg, l, *smp[b])
            scale = np.max(np.abs(ref))
            big = np.abs(ref) > 1e-200
            assert np.max(np.abs(U[b, l][big] - ref[big]) / np.abs(ref[big])) < 2e-13, (b, l)
            assert np.max(np.abs(U[b, l] - ref)) < 1e-13 * scale
    for l in (0, 3):
        assert relerr(U[0, l], g[f"s0_l{l}_U"], floor=1e-200) < 2e-13
    # the reference's literal form (z = 2 pi r r'/beta^2) on a small mesh where spherical_jn(-i z) is still finite
    rs = np.linspace(0.05, 2.0, 12)
    a_, b_ = np.meshgrid(rs, rs, indexing="ij")
    for ell in (0, 2, 5):
        z = 2 * np.pi * a_ * b_ / 1.1**2
        ref = np.real(np.exp(-(a_**2 + b_**2) / 1.1**2) * 2 * 1j**ell * z * sc.spherical_jn(ell, -1j * z) / (1.1 * np.sqrt(np.pi)))
        got = nonlocal_forms.perey_buck_nonlocal(rs, rs, 1.1, ell).cpu().numpy()
        assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-12
    # and the kernels feed the packed solver: S of the golden (sample 0, l = 0, 3)
    sys4 = ProjectileTargetSystem(float(g["a"]), 6, Ztarget=82, Zproj=0)
    channels, asym = sys4.get_partial_wave_channels(*g["kin"])
    for l in (0, 3):
        free = torch.tensor(solver.free_matrix(float(g["a"]), [l]), device="cuda")
        at = torch.tensor(np.array([[asym[l].Hp[0]], [asym[l].Hm[0]], [asym[l].Hpp[0]], [asym[l].Hmp[0]]]), device="cuda")
        out = solver.solve_batch(float(g["a"]), float(g["kin"][1]), k, 1, free, at, nonlocal_potential=torch.tensor(U[:1, l], device="cuda"))
        assert relerr(out["S"][0].cpu().numpy(), g[f"s0_l{l}_S"]) < RTOL_S


# ------------------------------------------------------------------ the reference's external-code regression, re-pointed
FRESCOX_CASES = ["F1", "F2", "F3", "F4", "F5", "F6", "F7", "F8"]


@pytest.mark.parametrize("cid", FRESCOX_CASES)
def test_frescox_regression_cases(golden, cid):
    """tests/regression/test_regression.py:9-20 of the reference, cases F1..F8 (Frescox, p / n + 78Ni, 6.9-200 MeV, nbasis 50-100,
    lmax 30-100): the same workspaces through this API, against Frescox within the reference's own tolerances (0.5-2 %) and
    against the reference's own numbers (S 1e-10, dsigma/dOmega 1e-8).  nbasis <= 64 takes the fused packed sweep, nbasis > 64
    the one-launch-per-energy batched path."""
    from jitr_b200.utils.kinematics import ChannelKinematics

    g = golden("regression_frescox")
    At, Zt, Ap, Zp = (int(v) for v in g[f"{cid}_AZ"])
    Rch, lmax, N = float(g[f"{cid}_match"][0]), int(g[f"{cid}_match"][1]), int(g[f"{cid}_match"][2])
    rx = ElasticReaction((At, Zt), (Ap, Zp), mass_target=float(g[f"{cid}_masses"][0]), mass_projectile=float(g[f"{cid}_masses"][1]))
    kin = ChannelKinematics(*[float(v) for v in g[f"{cid}_kin"]])
    ang = g[f"{cid}_angles"]
    ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, Rch, jitr_b200.Solver(N), lmax, ang)
    co = g[f"{cid}_coul"]
    t = lambda v: torch.tensor(v[None, :], device="cuda")  # noqa: E731
    x = ws.xs_batch(t(g[f"{cid}_central"]), t(g[f"{cid}_so"]), None if co.size == 0 else t(co))
    dsdo = x.dsdo[0].cpu().numpy()
    rtol, atol = (float(v) for v in g[f"{cid}_tol"])
    np.testing.assert_allclose(dsdo, g[f"{cid}_frescox"], rtol=rtol, atol=atol)           # the external code
    S, nl, status = ws.integral_workspace.smatrix_batch(t(g[f"{cid}_central"]), t(g[f"{cid}_so"]), None if co.size == 0 else t(co))
    L = int(g[f"{cid}_nl"])
    assert int(nl[0]) == L and int(status[0]) == 0
    Sg = S[0].cpu().numpy()
    assert relerr(Sg[0, :L], g[f"{cid}_splus"][:L]) < 2e-10 and relerr(Sg[1, 1:L], g[f"{cid}_sminus"][1:L]) < 2e-10
    assert np.all(Sg[:, L:] == 0)
    ref = g[f"{cid}_dsdo"]
    assert np.max(np.abs(dsdo - ref) / np.maximum(ref, 1e-6 * ref.max())) < 1e-7                     # the reference itself
    assert abs(float(x.rxn[0]) - float(g[f"{cid}_rxn"])) < RTOL_XS * abs(float(g[f"{cid}_rxn"]))


def test_indexed_free_matrices_equal_per_partial_wave_calls():
    """jitr_b200_solve_batched_indexed: one call for a batch that mixes partial waves (free matrix picked per system, asymptotics
    per system) gives bit for bit what one call per partial wave gives -- single channel (packed kernel) and coupled (CTA kernel)."""
    rng = np.random.default_rng(77)
    for N, nch in ((60, 1), (24, 3)):
        solver = jitr_b200.Solver(N)
        a, k0 = 9.0, 0.8
        E = list(12.0 - 0.5 * np.arange(nch))
        mu = [930.0] * nch
        ls = [0, 2, 5]
        B = 9
        free = torch.tensor(np.stack([solver.free_matrix(a, [l] * nch, E, mu) for l in ls]), device="cuda")
        asym = torch.tensor(np.stack([np.array([orc.coulomb_hankel(l, 0.0, a) for _ in range(nch)]).T for l in ls]), device="cuda")  # (3, 4, nch)
        r = solver.radial_grid(a, k0)
        loc = np.zeros((B, nch, nch, N), dtype=np.complex128)
        for i in range(nch):
            for j in range(i, nch):
                v = (-30.0 - 3j) * (i == j) / (1 + np.exp((r - 4.0) / 0.6)) * (1 + 0.1 * rng.standard_normal((B, 1))) + 2.0 * (i != j) * np.exp(-((r - 4) / 1.5) ** 2)
                loc[:, i, j] = v
                loc[:, j, i] = v
        loc_t = torch.tensor(loc, device="cuda")
        idx = torch.arange(3, dtype=torch.int32, device="cuda").repeat_interleave(B)
        one = solver.solve_batch(a, E[0], k0, nch, free, asym[idx.long()].contiguous(), local_potential=loc_t.repeat(3, 1, 1, 1), free_index=idx)
        assert int(one["status"].abs().sum()) == 0
        for q, l in enumerate(ls):
            per = solver.solve_batch(a, E[0], k0, nch, free[q], asym[q], local_potential=loc_t)
            assert torch.equal(one["S"][q * B:(q + 1) * B], per["S"]) and torch.equal(one["R"][q * B:(q + 1) * B], per["R"])


class _Guarded:
    """Output / input buffers cut out of a larger allocation whose margins hold a canary pattern: a kernel that writes out of
    bounds changes a margin, a kernel that reads out of bounds and uses the value sees NaNs (the device-side stand-in for a
    memory checker: bounds evidence from our own guard bands, small cases and bit-for-bit comparison with an unguarded run)."""

    MARGIN = 4096  # bytes on both sides

    def __init__(self):
        self.bufs = []

    def empty(self, shape, dtype, fill=None):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        pad = (-n) % 16
        raw = torch.full((self.MARGIN + n + pad + self.MARGIN,), 0xA5, dtype=torch.uint8, device="cuda")
        if dtype in (torch.float64, torch.complex128):  # margins that poison any value computed from them
            raw[: self.MARGIN].view(torch.float64).fill_(float("nan"))
            raw[self.MARGIN + n + pad:].view(torch.float64).fill_(float("nan"))
        view = raw[self.MARGIN: self.MARGIN + n].view(dtype).reshape(shape)
        if fill is not None:
            view.copy_(fill)
        self.bufs.append((raw, n + pad, raw[: self.MARGIN].clone(), raw[self.MARGIN + n + pad:].clone()))
        return view

    def check(self):
        torch.cuda.synchronize()
        for raw, n, lo, hi in self.bufs:
            assert torch.equal(raw[: self.MARGIN], lo) or bool(torch.equal(raw[: self.MARGIN].view(torch.int64), lo.view(torch.int64))), "write below a buffer"
            assert torch.equal(raw[self.MARGIN + n:].view(torch.int64), hi.view(torch.int64)), "write beyond a buffer"


@pytest.mark.parametrize("N,lmax,B", [(12, 9, 37), (40, 30, 75), (40, 30, 3), (56, 24, 41), (64, 40, 13)])
def test_guard_bands_around_sweep_buffers(golden, N, lmax, B):
    """Every path of the fused sweep (small modes, the N = 40 symmetric mode incl. the partial-wave split of small batches,
    the packed modes), odd batch sizes: inputs and outputs live between NaN / canary guard bands; the results must equal the
    unguarded run bit for bit and the bands must be untouched."""
    g = golden("kduq_params")
    rows = np.zeros((B, 40))
    rows[:, :37] = np.tile(g["federal_n"], (B // len(g["federal_n"]) + 1, 1))[:B]
    rx = ElasticReaction((208, 82), (1, 0), mass_target=193687.12278341982, mass_projectile=939.5654983938299)
    solver = jitr_b200.Solver(N)
    ang = np.linspace(0.1, np.pi, 23)
    wss = [jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(E), 12.0, solver, lmax, ang) for E in (7.0, 33.0, 61.0)]
    sweep = jitr_b200.ElasticSweep(wss)
    model = kduq.KDUQ((1, 0))
    p = torch.tensor(rows, device="cuda")
    S0, nl0, st0 = sweep.smatrix_params(model, p)
    x0 = sweep.observables(S0, nl0)
    G = _Guarded()
    pg = G.empty(p.shape, torch.float64, fill=p)
    out = (G.empty(S0.shape, torch.complex128), G.empty(nl0.shape, nl0.dtype), G.empty(st0.shape, st0.dtype))
    S, nl, st = sweep.smatrix_params(model, pg, out=out)
    obs = (G.empty(x0.dsdo.shape, torch.float64), G.empty(x0.dsdo.shape, torch.float64), G.empty(x0.dsdo.shape, torch.float64),
           G.empty((B, 3, 2), torch.float64))
    x = sweep.observables(S, nl, out=obs)
    G.check()
    assert torch.equal(nl, nl0) and torch.equal(st, st0)
    assert torch.equal(torch.view_as_real(S), torch.view_as_real(S0))
    for a, b in ((x.dsdo, x0.dsdo), (x.Ay, x0.Ay), (x.Q, x0.Q), (x.t, x0.t), (x.rxn, x0.rxn)):
        assert torch.equal(a, b)


@pytest.mark.parametrize("N,nch,B", [(60, 1, 19), (23, 1, 7), (50, 3, 5), (21, 2, 9)])
def test_guard_bands_around_general_solver_buffers(N, nch, B):
    """solve_batched (packed single-channel kernel and the blocked multi-channel kernel): every input between NaN guard bands
    -- an over-read that reaches the arithmetic poisons the result -- compared bit for bit with the unguarded call."""
    rng = np.random.default_rng(N * 10 + nch)
    solver = jitr_b200.Solver(N)
    a, k, E = 11.0, 0.9, 14.0
    free = torch.tensor(solver.free_matrix(a, [1] * nch), device="cuda")
    H = np.array([orc.coulomb_hankel(1, 0.0, a) for _ in range(nch)]).T
    asym = torch.tensor(np.ascontiguousarray(H), device="cuda")
    r = solver.radial_grid(a, k)
    V = np.zeros((B, nch, nch, N), dtype=np.complex128)
    for b in range(B):
        for i in range(nch):
            for j in range(i, nch):
                v = (-40.0 - 5j if i == j else 3.0) * (1 + 0.1 * rng.standard_normal()) / (1 + np.exp((r - 4.0) / 0.6))
                V[b, i, j] = V[b, j, i] = v
    Vt = torch.tensor(V, device="cuda")
    o0 = solver.solve_batch(a, E, k, nch, free, asym, local_potential=Vt)
    G = _Guarded()
    Vg = G.empty(Vt.shape, torch.complex128, fill=Vt)
    fg = G.empty(free.shape, torch.complex128, fill=free)
    ag = G.empty(asym.shape, torch.complex128, fill=asym)
    o1 = solver.solve_batch(a, E, k, nch, fg, ag, local_potential=Vg)
    G.check()
    assert int(o0["status"].abs().sum()) == 0
    for key in ("R", "S", "uext"):
        if key in o0:
            assert torch.equal(torch.view_as_real(o0[key]), torch.view_as_real(o1[key])), key
