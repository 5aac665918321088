"""Pin the CPU oracle (oracle/jitr_oracle.py) against golden vectors produced by
the real reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import jitr_oracle as orc

RTOL_S = 1e-11   # oracle vs reference S-matrix (same algorithm, same LAPACK): rounding-level
RTOL_XS = 1e-10


def normerr(a, b):
    """max-norm relative error (for vectors whose small entries carry no relative accuracy)."""
    a = np.asarray(a)
    b = np.asarray(b)
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


ELASTIC = [
    ("config1_n208Pb_30MeV", "kduq"), ("config3_p48Ca_30MeV", "kduq"),
    ("config2_n40Ca_E00", "kduq"), ("config2_n40Ca_E31", "kduq"),
    ("config2_n208Pb_E06", "kduq"), ("config2_n208Pb_E15", "kduq"),
    ("chuq_p90Zr_25MeV", "chuq"), ("chuq_n90Zr_12MeV", "chuq"), ("localomp_p54Fe_35MeV", "local"),
    ("wlh_n208Pb_14MeV", "wlh"), ("wlh_p48Ca_45MeV", "wlh"),
]


def _potentials(g, model, p):
    tgt = tuple(int(v) for v in g["target"])
    prj = tuple(int(v) for v in g["projectile"])
    Elab = float(g["kin"][0])
    r = g["rgrid"]
    if model == "kduq":
        return orc.kduq_potentials(r, prj, tgt, Elab, p)
    if model == "chuq":
        return orc.chuq_potentials(r, prj, tgt, Elab, p)
    if model == "wlh":
        return orc.wlh_potentials(r, prj, tgt, Elab, p)
    return orc.local_omp_potentials(r, tgt[0], tgt[1] * prj[1], p)


@pytest.mark.parametrize("name,model", ELASTIC)
def test_elastic_workspace(golden, name, model):
    g = golden(name)
    N, lmax = int(g["N"]), int(g["lmax"])
    kin = tuple(g["kin"])
    tgt, prj = g["target"], g["projectile"]
    # kinematics restatement
    kin2 = orc.semi_relativistic_kinematics(g["masses"][0], g["masses"][1], kin[0], int(tgt[1] * prj[1]))
    np.testing.assert_allclose(kin2, kin, rtol=1e-14)
    ws = orc.ElasticWorkspace(N, lmax, float(g["Rch"]), kin, g["angles"], int(tgt[1] * prj[1]))
    # setup tables
    assert relerr(ws.asym, g["asym"]) < 1e-13
    np.testing.assert_allclose(ws.rgrid, g["rgrid"], rtol=1e-14)
    np.testing.assert_allclose(ws.free[0], g["free0"], rtol=1e-13)
    np.testing.assert_allclose(ws.free[3], g["free3"], rtol=1e-13)
    np.testing.assert_allclose(ws.b, g["basis_boundary"], rtol=1e-13)
    np.testing.assert_allclose(ws.P, g["P"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(ws.f_c, g["f_c"], rtol=1e-13)
    nsamp = min(len(g["params"]), 6)
    for i in range(nsamp):
        c, so, co = _potentials(g, model, g["params"][i])
        np.testing.assert_allclose(c, g["central"][i], rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(so, g["so"][i], rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(co, g["coul"][i], rtol=1e-13, atol=1e-300)
        sp, sm = ws.smatrix(c, so, co)
        nl = int(g["nl"][i])
        assert len(sp) == nl
        assert relerr(sp, g["splus"][i, :nl]) < RTOL_S
        assert relerr(sm[1:], g["sminus"][i, 1:nl]) < RTOL_S
        dsdo, Ay, Q, tot, rxn = ws.xs(c, so, co)
        assert relerr(dsdo, g["dsdo"][i]) < RTOL_XS
        np.testing.assert_allclose(Ay, g["Ay"][i], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(Q, g["Q"][i], rtol=1e-8, atol=1e-10)
        assert abs(tot - g["tot"][i]) / g["tot"][i] < RTOL_XS
        assert abs(rxn - g["rxn"][i]) / g["rxn"][i] < RTOL_XS


def test_kduq_shape_params(golden):
    g = golden("kduq_shape_params")
    P = golden("kduq_params")
    for key in g.files:
        tag, A, Z, E = key.split("_")
        prj = (1, 0) if tag == "n" else (1, 1)
        samples = P["federal_n"] if tag == "n" else P["federal_p"]
        for i in range(0, 416, 7):
            with np.errstate(over="ignore"):
                c, s, co = orc.kduq_shape_params(prj, (int(A), int(Z)), float(E), samples[i])
            np.testing.assert_allclose(list(c) + list(s) + list(co), g[key][i], rtol=1e-14)


def test_wlh_shape_params(golden):
    """wlh.calculate_params restated (both branches of the surface term, neutrons and protons)."""
    g = golden("wlh_shape_params")
    pr = golden("wlh_params")
    k = 0
    for prj_a, prj_z, A, Z, E in g["cases"]:
        prj, tgt = (int(prj_a), int(prj_z)), (int(A), int(Z))
        plist = [pr["mean_n"]] + list(pr["samples_n"][:3]) if prj == (1, 0) else [pr["mean_p"]] + list(pr["samples_p"][:3])
        for p in plist:
            c, s_, co = orc.wlh_shape_params(prj, tgt, float(E), p)
            np.testing.assert_allclose(np.array(list(c) + list(s_) + list(co), dtype=np.float64), g["shapes"][k], rtol=1e-14, atol=0)
            k += 1
    assert k == len(g["shapes"])


def test_local_rk_case(golden):
    """restated tests/test_local.py: N=70, a=10 pi, WS+Coulomb, l = 0..2 x 10 energies."""
    g = golden("test_local")
    N, a = int(g["N"]), float(g["a"])
    for i in range(10):
        Elab, Ecm, mu, k, eta = g["kin"][i]
        for l in range(3):
            asym = tuple(np.array([v]) for v in orc.coulomb_hankel(l, eta, a))
            assert relerr(np.array(asym).ravel(), g["asym"][i, l]) < 1e-13
            R, S, u = orc.solver_solve(N, a, [l], [Ecm], [k], [mu], asym, local_potential=g["pot"][i])
            assert relerr(S[0, 0], g["S"][i, l]) < RTOL_S
            assert relerr(R[0, 0], g["R"][i, l]) < 1e-10
            assert relerr(u[0], g["uext"][i, l]) < 1e-10


def test_wavefunction_case(golden):
    g = golden("test_wavefunction")
    N, a = int(g["N"]), float(g["a"])
    Elab, Ecm, mu, k, eta = g["kin"]
    for l in (0, 1, 4):
        asym = tuple(np.array([v]) for v in g[f"asym{l}"])
        R, S, x, u = orc.solver_solve(N, a, [l], [Ecm], [k], [mu], asym, local_potential=g[f"pot{l}"], wavefunction=True)
        # N=100, strongly absorbed (|S| ~ 1e-2): the reference's own inverse-vs-solve noise floor is ~1e-10 here
        assert relerr(S, g[f"S{l}"]) < 1e-9
        assert np.max(np.abs(S - g[f"S{l}"])) < 1e-11
        assert normerr(x, g[f"x{l}"]) < 1e-10
        uint = sum(x[0, n] / a * orc.basis_function(N, n + 1, a, g["svals"]) for n in range(N))
        assert normerr(uint, g[f"uint{l}"]) < 1e-10


def test_coupled_config5(golden):
    g = golden("config5_coupled")
    for N in (40, 50):
        for l in range(4):
            asym = tuple(g[f"l{l}_asym"])
            R, S, x, u = orc.solver_solve(N, float(g["a"]), [l] * 3, g["E"], g["k"], g["mu"], asym,
                                          local_potential=g[f"N{N}_l{l}_V"], wavefunction=True)
            assert relerr(S, g[f"N{N}_l{l}_S"]) < 1e-10
            assert relerr(R, g[f"N{N}_l{l}_R"]) < 1e-9
            assert normerr(x, g[f"N{N}_l{l}_x"]) < 1e-10


def test_coupled_single(golden):
    g = golden("test_coupled_single")
    Elab, Ecm, mu, k, eta = g["kin"]
    asym = tuple(g["asym"])
    Rm, Sm, um = orc.solver_solve(40, float(g["a"]), [0, 0], [Ecm] * 2, [k] * 2, [mu] * 2, asym, local_potential=g["V2"])
    assert np.max(np.abs(Sm - g["Sm"])) < 1e-11
    asym1 = tuple(v[:1] for v in asym)
    R1, S1, u1 = orc.solver_solve(40, float(g["a"]), [0], [Ecm], [k], [mu], asym1, local_potential=g["V1"])
    assert relerr(S1, g["S1"]) < RTOL_S
    np.testing.assert_almost_equal(Sm[0, 0], S1[0, 0])


def test_nonlocal(golden):
    g = golden("nonlocal_yamaguchi")
    Elab, Ecm, mu, k, eta = g["kin"]
    asym = tuple(np.array([v]) for v in g["yam_asym"])
    for N in (20, 40):
        R, S, u = orc.solver_solve(N, float(g["a"]), [0], [Ecm], [k], [mu], asym, nonlocal_potential=g[f"yam{N}_V"])
        assert relerr(S, g[f"yam{N}_S"]) < RTOL_S
    delta = np.rad2deg(np.real(np.log(S[0, 0]) / 2j))
    assert abs(delta - (-66.98635852)) < 1e-6  # SURVEY appendix B (mesh value, not the analytic one)
    g = golden("config4_nonlocal")
    Elab, Ecm, mu, k, eta = g["kin"]
    for l in (0, 3):
        asym = tuple(np.array([v]) for v in g[f"l{l}_asym"])
        R, S, u = orc.solver_solve(60, float(g["a"]), [l], [Ecm], [k], [mu], asym, nonlocal_potential=g[f"s0_l{l}_U"])
        assert relerr(S, g[f"s0_l{l}_S"]) < RTOL_S


def test_survey_anchor_values(golden):
    """SURVEY.md appendix B sanity anchors (captured independently in the survey session)."""
    g = golden("config1_n208Pb_30MeV")
    assert int(g["nl"][0]) == 21
    assert abs(g["splus"][0, 0] - (-0.4133998941888191 + 0.08618832270068263j)) < 1e-12
    assert abs(g["tot"][0] - 5045.3932311612) < 1e-6
    assert abs(g["rxn"][0] - 2418.2642857019428) < 1e-6
    assert abs(g["dsdo"][0][90] - 8.6793242520790077) < 1e-8
    g = golden("nb_kduq_uq_demo")
    assert round(float(g["pct84_idx9"]), 6) == 1.163423  # kduq_uq_demo.ipynb cell 20


@pytest.mark.parametrize("name", ["qe_reftest", "qe_48Ca_35MeV", "qe_48Ca_35MeV_nobreak"])
def test_quasielastic_pn_workspace(golden, name):
    """Restatement of xs/quasielastic_pn.py against the reference's own outputs (T_lj, S of both channels,
    dsigma/dOmega; early break on and off)."""
    g = golden(name)
    AZ = g["AZ"]
    kx = orc.exit_kinematics(g["kin_entrance"], g["masses"], float(g["Q"]), float(g["Ex"]), 0.0, Zz_exit=int(AZ[5] * AZ[7]))
    assert np.allclose(np.array(kx), g["kin_exit"], rtol=1e-14, atol=0)
    w = orc.QuasielasticWorkspace(int(g["N"]), int(g["lmax"]), float(g["Rch"]), g["kin_entrance"], g["kin_exit"], g["angles"],
                                  int(AZ[0]), int(AZ[1]), int(AZ[1] * AZ[3]), 0, float(g["tol"]))
    gf = g["geometric_factor"]
    assert np.abs(w.geometric_factor - gf).max() <= 1e-14 * np.abs(gf).max()
    assert abs(w.xs_factor - g["xs_factor"]) <= 1e-15 * g["xs_factor"]
    assert abs(w.isovector_factor - g["isovector_factor"]) <= 1e-15
    for b in range(min(3, g["U"].shape[0])):
        T, Sn, Sp = w.tmatrix(*g["U"][b])
        assert np.abs(T - g["Tpn"][b]).max() <= 1e-11 * np.abs(g["Tpn"][b]).max()
        assert np.abs(Sn - g["Sn"][b]).max() <= 1e-11
        assert np.abs(Sp - g["Sp"][b]).max() <= 1e-11
        assert np.abs(w.xs(*g["U"][b]) / g["xs"][b] - 1).max() <= 1e-10


def test_dom_parameter_map_and_forms(golden):
    """Dispersive optical model: depth functions, analytic surface dispersion integral (both ws2 branches) and the
    radial forms against the reference's dom.calculate_params / central / spin_orbit."""
    g = golden("dom_params")
    k = 0
    for case in g["cases"]:
        prj, tgt, Ecm, Ef = (int(case[0]), int(case[1])), (int(case[2]), int(case[3])), float(case[4]), float(case[5])
        for p in g["params"]:
            c, s, co = orc.dom_shape_params(prj, tgt, Ecm, Ef, p)
            got, ref = np.array(list(c) + list(s) + list(co)), g["shapes"][k]
            assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-12)) < 1e-12
            k += 1
    c, so, _ = orc.dom_potentials(g["r"], (1, 1), (48, 20), 44.0, -8.4, g["params"][2])
    assert np.max(np.abs(c - g["central"])) < 1e-12 * np.max(np.abs(g["central"]))
    assert np.max(np.abs(so - g["spin_orbit"])) < 1e-12 * np.max(np.abs(g["spin_orbit"]))
    got = np.array([orc.dom_delta_Vs(x, 15.0, 12.5, w2) for x in (-40.0, -3.0, 0.0, 0.7, 12.5, 80.0) for w2 in (0.0, 0.021)])
    assert np.allclose(got, g["dVs"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("name", ["dom_n40Ca_14MeV", "dom_p208Pb_30MeV"])
def test_dom_elastic_workspace(golden, name):
    g = golden(name)
    prj, tgt = tuple(int(v) for v in g["projectile"]), tuple(int(v) for v in g["target"])
    ow = orc.ElasticWorkspace(int(g["N"]), int(g["lmax"]), float(g["Rch"]), tuple(g["kin"]), g["angles"], prj[1] * tgt[1], asym_table=g["asym"])
    for i in (0, 1, 4):
        c, so, co = orc.dom_potentials(g["rgrid"], prj, tgt, float(g["kin"][1]), float(g["Ef"]), g["params"][i])
        assert np.max(np.abs(c - g["central"][i])) < 1e-11 * np.max(np.abs(g["central"][i]))
        sp, sm = ow.smatrix(c, so, co)
        assert len(sp) == int(g["nl"][i])
        assert np.max(np.abs(sp - g["Splus"][i, : len(sp)])) < 1e-10
        dsdo = ow.xs(c, so, co)[0]
        assert np.max(np.abs(dsdo / g["dsdo"][i] - 1)) < 1e-9


def test_parity_sweep_runner_vs_reference_golden(golden):
    """oracle/parity_sweep.py (the multi-process evaluator behind the large-sample parity gate) reproduces the
    reference's config-2 goldens on a two-workspace table."""
    from oracle import parity_sweep

    names = ("config2_n40Ca_E06", "config2_n208Pb_E15")
    gs = [golden(n) for n in names]
    tab = [((int(g["target"][0]), int(g["target"][1])), float(g["kin"][0]), float(g["Rch"]), tuple(float(v) for v in g["kin"])) for g in gs]
    rows = gs[0]["params"][:4]
    assert np.array_equal(rows, gs[1]["params"][:4])
    o = parity_sweep.run(tab, rows, 40, 30, angles=gs[0]["angles"], procs=2)
    for e, g in enumerate(gs):
        for i in range(4):
            L = int(g["nl"][i])
            assert int(o["nl"][i, e]) == L
            assert np.max(np.abs(o["splus"][i, e, :L] - g["splus"][i, :L]) / np.abs(g["splus"][i, :L])) < 1e-11
            assert np.max(np.abs(o["sminus"][i, e, 1:L] - g["sminus"][i, 1:L]) / np.abs(g["sminus"][i, 1:L])) < 1e-11
            assert np.all(o["splus"][i, e, L:] == 0)
            assert np.max(np.abs(o["dsdo"][i, e] - g["dsdo"][i]) / g["dsdo"][i]) < 1e-10
            assert abs(o["tot_rxn"][i, e, 1] - g["rxn"][i]) / g["rxn"][i] < 1e-11


def test_refinement_referee_matches_40_digit_solve():
    """oracle/mp_truth.py: the fast referee (float64 LU + 80-bit iterative refinement) agrees with the mpmath 40-digit
    solve of the same float64 matrix to 1e-13 -- three orders below the 1e-10 parity gate it arbitrates."""
    from jitr_b200 import workloads
    from oracle import mp_truth

    tab = workloads.energy_table()
    rows = workloads.synthetic_kduq_samples(8)
    items = [(0, 5, 3, 0), (1, 40, 7, 1), (5, 0, 0, 0), (7, 63, 12, 1)]
    exact = mp_truth.referee(tab, rows, items, dps=40, procs=2)
    fast = mp_truth.referee(tab, rows, items, dps=0, procs=2)
    for a, b in zip(exact, fast):
        assert abs(a - b) / abs(a) < 1e-13


def test_one_ulp_input_noise_of_the_referee():
    """oracle/mp_truth.smatrix_input_noise: the change one ulp of uncertainty in the matrix entries makes to an S-matrix element.
    It must be a small multiple of machine epsilon for well-conditioned elements, it must reproduce (seeded), and the float64
    oracle itself must sit within a few of these 'input ulps' of the exact value -- the yardstick the large-sample gate uses
    for the handful of elements of 10^7 that no float64 algorithm pins to 1e-10."""
    from jitr_b200 import workloads
    from oracle import mp_truth, parity_sweep

    tab = workloads.energy_table()
    rows = workloads.synthetic_kduq_samples(8)
    items = [(0, 5, 3, 0), (1, 40, 7, 1), (5, 0, 0, 0), (7, 63, 12, 1)]
    noise = mp_truth.referee(tab, rows, items, dps=-1, procs=2)
    again = mp_truth.referee(tab, rows, items, dps=-1, procs=1)
    assert noise == again
    exact = mp_truth.referee(tab, rows, items, dps=0, procs=2)
    o = parity_sweep.run([tab[e] for e in sorted({e for _, e, _, _ in items})], rows, procs=2)
    emap = {e: k for k, e in enumerate(sorted({e for _, e, _, _ in items}))}
    for (i, e, l, jj), nz, st in zip(items, noise, exact):
        assert 1e-17 < nz < 1e-9
        so_ = (o["splus"] if jj == 0 else o["sminus"])[i, emap[e], l]
        assert abs(so_ - st) / abs(st) < max(1e-12, 16 * nz)
