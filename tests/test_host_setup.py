"""Host-side mirror of the reference interface (setup math, classes, error behaviour) vs golden vectors.  CPU only."""
import numpy as np
import pytest

import jitr_b200
from jitr_b200 import reactions, utils
from jitr_b200.free_solutions import coulomb_hankel_table
from jitr_b200.optical_potentials import chuq, kduq
from jitr_b200.reactions import ElasticReaction, ProjectileTargetSystem


def test_kinematics_and_tables(golden):
    for name in ("config1_n208Pb_30MeV", "config3_p48Ca_30MeV"):
        g = golden(name)
        rx = ElasticReaction(tuple(g["target"]), tuple(g["projectile"]), mass_target=g["masses"][0], mass_projectile=g["masses"][1])
        kin = rx.kinematics(float(g["kin"][0]))
        np.testing.assert_allclose(list(kin), g["kin"], rtol=1e-15)
        N, lmax = int(g["N"]), int(g["lmax"])
        solver = jitr_b200.Solver(N)
        ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, float(g["Rch"]), solver, lmax, g["angles"])
        iw = ws.integral_workspace
        np.testing.assert_allclose(ws.radial_grid(), g["rgrid"], rtol=1e-15)
        np.testing.assert_allclose(iw.free_matrices[0], g["free0"], rtol=1e-13)
        np.testing.assert_allclose(iw.free_matrices[3], g["free3"], rtol=1e-13)
        np.testing.assert_allclose(iw.basis_boundary, g["basis_boundary"], rtol=1e-14)
        assert np.max(np.abs(iw._asym_table - g["asym"]) / np.abs(g["asym"])) < 1e-13
        np.testing.assert_allclose(ws.P_l_costheta, g["P"], rtol=1e-15, atol=0)
        np.testing.assert_allclose(ws.P_1_l_costheta, g["P1"], rtol=1e-15, atol=0)
        np.testing.assert_allclose(ws.sigma_l, g["sigma_l"], rtol=1e-15)
        np.testing.assert_allclose(ws.f_c, g["f_c"], rtol=1e-14)
        if ws.rutherford is not None:
            np.testing.assert_allclose(ws.rutherford, g["rutherford"], rtol=1e-14)
        # channel objects keep the reference fields
        ch = iw.channels[3][1]
        assert ch.size == 1 and ch.l[0] == 3 and ch.couplings[0, 0] == -4.0
        assert iw.l_dot_s.shape == (lmax, 2)


def test_system_classes(golden):
    g = golden("config5_coupled")
    cm = g["coupling"]
    sysc = ProjectileTargetSystem(
        channel_radius=float(g["a"]), lmax=3, mass_target=1.0, mass_projectile=1.0, Ztarget=20, Zproj=1,
        channel_levels=np.array([0, 2.3, 3.1]), coupling=lambda l: cm,
    )
    channels, asym = sysc.get_partial_wave_channels(5.0, g["E"], g["mu"], g["k"], g["eta"])
    for l in range(4):
        ref = g[f"l{l}_asym"]
        got = np.array([asym[l].Hp, asym[l].Hm, asym[l].Hpp, asym[l].Hmp])
        assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-13
        assert channels[l].size == 3 and channels[l].num_channels == 3
    with pytest.raises(AssertionError):
        reactions.Channels(np.ones(2), np.ones(2), np.ones(2), np.ones(2), 1.0, np.ones(2), np.eye(3))


def test_solver_host_helpers_and_errors():
    s = jitr_b200.Solver(12)
    F = s.free_matrix(7.0, [0, 2], E=[5.0, 3.0], mu=[900.0, 910.0])
    assert F.shape == (24, 24)
    np.testing.assert_allclose(s.get_channel_block(F, 0, 1), 0)
    np.testing.assert_allclose(s.get_channel_block(F, 1), s.kernel.quadrature.kinetic_matrix(7.0, 2) * 900.0 / 910.0 - np.eye(12) * 3.0 / 5.0)
    # element-wise definition agrees with the vectorised That
    q = s.kernel.quadrature
    T = q.kinetic_matrix(7.0, 2)
    for n, m in ((1, 1), (2, 5), (12, 3)):
        assert abs(T[n - 1, m - 1] - q.kinetic_operator_element(n, m, 7.0, 2)) < 1e-12 * abs(T[n - 1, m - 1])
    with pytest.raises(ValueError):
        s.interaction_matrix(1.0, 1.0, 7.0, 2, local_potential=np.zeros(12))
    with pytest.raises(ValueError):
        s.interaction_matrix(1.0, 1.0, 7.0, 1, local_potential=np.zeros(11))
    with pytest.raises(ValueError):
        s.interaction_matrix(1.0, 1.0, 7.0, 1, nonlocal_potential=np.zeros((12, 11)))
    with pytest.raises(NotImplementedError):
        jitr_b200.Solver(10, "Laguerre")
    with pytest.raises(ValueError):
        jitr_b200.xs.elastic.check_angles(np.array([[0.1, 0.2]]))
    with pytest.raises(ValueError):
        jitr_b200.IntegralWorkspace(reactions.Reaction((40, 20), (1, 0), process="abs"), None, 10.0, s, 3)


def test_interaction_matrix_matches_oracle():
    from oracle import jitr_oracle as orc

    s = jitr_b200.Solver(10)
    rng = np.random.default_rng(1)
    loc = rng.standard_normal((2, 2, 10)) + 1j * rng.standard_normal((2, 2, 10))
    nl = rng.standard_normal((2, 2, 10, 10)) + 1j * rng.standard_normal((2, 2, 10, 10))
    got = s.interaction_matrix(1.3, 12.0, 9.0, 2, local_potential=loc, nonlocal_potential=nl)
    ref = orc.interaction_matrix(10, 1.3, 12.0, 9.0, 2, loc, nl)
    np.testing.assert_allclose(got, ref, rtol=1e-14)


def test_param_vectors(golden):
    g = golden("kduq_params")
    np.testing.assert_array_equal(kduq.get_kd03((1, 0)), g["kd03_n"])
    np.testing.assert_array_equal(kduq.get_kd03((1, 1)), g["kd03_p"])
    assert kduq.get_param_names((1, 0)) == list(g["names_n"])
    assert kduq.get_param_names((1, 1)) == list(g["names_p"])
    c = golden("chuq_params")
    np.testing.assert_array_equal(chuq.get_ch89(), c["ch89"])
    assert chuq.get_param_names() == list(c["names"])
    w = golden("wlh_params")
    from jitr_b200.optical_potentials import wlh

    np.testing.assert_array_equal(wlh.get_wlh_mean((1, 0)), w["mean_n"])
    np.testing.assert_array_equal(wlh.get_wlh_mean((1, 1)), w["mean_p"])
    assert wlh.get_param_names((1, 0)) == list(w["names_n"]) and wlh.get_param_names((1, 1)) == list(w["names_p"])
    with pytest.raises(RuntimeError):
        kduq.KDUQ((4, 2))
    with pytest.raises(RuntimeError):
        wlh.WLH((2, 1))


def test_neutral_asymptotics_closed_form_matches_mpmath():
    from oracle import jitr_oracle as orc

    for a in (3.3, 14.481215395437799, 31.4):
        tab = coulomb_hankel_table(12, 0.0, a)
        ref = np.array([orc.coulomb_hankel(l, 0.0, a) for l in range(13)])
        assert np.max(np.abs(tab - ref) / np.abs(ref)) < 5e-13


def test_utils():
    assert utils.suggested_basis_size(10 * np.pi) == 50
    assert abs(utils.interaction_range(208) - 1.2 * 208 ** (1 / 3)) < 1e-15
    k = utils.classical_kinematics(44657.0, 938.3, 42.1, 20)
    assert abs(k.Ecm - 44657.0 / (44657.0 + 938.3) * 42.1) < 1e-12


def test_quasielastic_workspace_host_setup(golden):
    """Host side of xs/quasielastic_pn.py:83-217 (no GPU): exit kinematics, isovector factor, kinematic
    factor and the geometric factors (closed-form Clebsch-Gordan instead of sympy) against the reference's."""
    from jitr_b200.xs.quasielastic_pn import Workspace

    g = golden("qe_48Ca_35MeV")
    AZ, m = g["AZ"], g["masses"]
    rx = reactions.Reaction((int(AZ[0]), int(AZ[1])), (int(AZ[2]), int(AZ[3])), (int(AZ[4]), int(AZ[5])), (int(AZ[6]), int(AZ[7])),
                            mass_target=m[0], mass_projectile=m[1], mass_product=m[2], mass_residual=m[3])
    assert abs(rx.Q - float(g["Q"])) < 1e-9
    ke = rx.kinematics(float(g["Elab"]))
    kx = rx.kinematics_exit(ke, float(g["Ex"]))
    assert np.allclose(np.array(tuple(ke)), g["kin_entrance"], rtol=1e-13)
    assert np.allclose(np.array(tuple(kx)), g["kin_exit"], rtol=1e-12)
    ws = Workspace(rx, ke, kx, jitr_b200.Solver(int(g["N"])), g["angles"], int(g["lmax"]), float(g["Rch"]))
    gf = g["geometric_factor"]
    assert np.abs(ws.geometric_factor - gf).max() <= 1e-12 * np.abs(gf).max()
    assert abs(ws.xs_factor / g["xs_factor"] - 1) < 1e-12
    assert abs(ws.isovector_factor - g["isovector_factor"]) < 1e-15
    assert np.allclose(ws.radial_grid(), g["rgrid"], rtol=1e-12)
    with pytest.raises(ValueError):
        Workspace(reactions.Reaction((48, 20), (1, 1)), ke, kx, jitr_b200.Solver(8), g["angles"], 3, 8.0)
    with pytest.raises(ValueError):
        reactions.Reaction((48, 20), (1, 1)).kinematics_exit(ke)


def test_steed_coulomb_functions_match_mpmath():
    """(f)2: the Coulomb-Hankel table off mpmath.  Steed's continued fractions against mpmath.coulombf/g (the
    reference's source, utils/free_solutions.py:39-50) at and above the l = 0 turning point rho = 2 eta -- also at the nodes
    of F_0; inside it the table comes from the inward-integration route (next test), so it is accurate everywhere."""
    import mpmath

    from jitr_b200 import free_solutions as fs

    worst = 0.0
    for eta, rho in [(0.05, 1.0), (0.59, 14.2), (1.3, 6.0), (3.0, 10.0), (5.6, 15.0), (9.0, 18.5), (15.0, 40.0),
                     (0.7553302787187184, 8.704174551241188)]:  # the last one sits on a node of F_0 (|G_0/F_0| = 166)
        assert rho >= 2 * eta
        F, G, Fp, Gp = fs.coulomb_fg_steed(30, eta, rho)
        for l in (0, 1, 7, 18, 30):
            Fm, Gm = float(mpmath.coulombf(l, eta, rho)), float(mpmath.coulombg(l, eta, rho))
            # relative to |H| = |G + iF| (G_l crosses zero in l at fixed rho: its own relative error means nothing there)
            # (and F_0 has nodes in rho: at one of them its own relative error loses |G_0/F_0| ulps, harmlessly)
            worst = max(worst, abs((G[l] - Gm) + 1j * (F[l] - Fm)) / abs(Gm + 1j * Fm), abs(F[l] / Fm - 1) * min(1.0, abs(Fm / Gm)))
            # Wronskian F' G - F G' = 1
            assert abs(Fp[l] * G[l] - F[l] * Gp[l] - 1) < 1e-11
    assert worst < 2e-14
    # table = the reference's definition (mpmath values, derivative by the l -> l+1 recurrence), both regimes
    from oracle import jitr_oracle as orc

    for eta, s in [(0.5903268765300718, 14.235372770878110), (2.2, 9.0), (6.0, 4.0)]:  # the last one is sub-barrier
        fs._TABLE_CACHE.clear()
        tab = fs.coulomb_hankel_table(12, eta, s)
        for l in (0, 3, 12):
            ref = np.array(orc.coulomb_hankel(l, eta, s))
            assert np.max(np.abs(tab[l] - ref) / np.abs(ref)) < 2e-13


def test_dom_host_parameter_rows(golden):
    """optical_potentials/dom.py on the host: calculate_params and the vectorised per-(sample, energy) canonical rows the
    device consumes, against the reference's calculate_params."""
    from jitr_b200.optical_potentials import dom

    g = golden("dom_params")
    assert dom.get_param_names() == [str(n) for n in g["names"]]
    k = 0
    for case in g["cases"]:
        prj, tgt, Ecm, Ef = (int(case[0]), int(case[1])), (int(case[2]), int(case[3])), float(case[4]), float(case[5])
        rows = dom.shape_rows(prj, tgt, np.array([Ecm, Ecm + 1.0]), Ef, g["params"])  # (B, 2, 18)
        for b, p in enumerate(g["params"]):
            c, s, co = dom.calculate_params(prj, tgt, Ecm, Ef, *p)
            ref = g["shapes"][k]
            got = np.array(list(c) + list(s) + list(co))
            assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-12)) < 1e-12
            V, R0, av, Wv, dVv, Rw, aw, Ws, dVs, Rd, ad, Vso, Wso, Rso, aso, zz, RC = ref
            want = np.array([V, R0, av, Wv, Rw, aw, dVs, Ws, Rd, ad, 2 * Vso, Rso, aso, 2 * Wso, Rso, aso, RC, dVv])
            assert np.max(np.abs(rows[b, 0] - want) / np.maximum(np.abs(want), 1e-12)) < 1e-12
            k += 1
    with pytest.raises(ValueError):
        dom.shape_rows((1, 0), (40, 20), np.array([10.0]), -8.0, np.zeros((2, 5)))


def test_coulomb_functions_below_turning_point_without_mpmath(monkeypatch):
    """utils/free_solutions.py:39-109 without multi-precision arithmetic in EVERY regime: inside the l = 0 turning point
    (rho < 2 eta, low-energy protons) G_0 is integrated inward from a point outside and F_0 follows from the Wronskian
    (free_solutions.coulomb_fg_below_turning_point): <= 1e-13 of mpmath on an (eta, rho, l) grid, and the table builder no
    longer reaches mpmath there."""
    import mpmath as mp

    from jitr_b200 import free_solutions as fs

    worst = 0.0
    for eta in (0.3, 0.59, 1.4, 3.0, 7.5, 15.0, 30.0):
        for frac in (0.02, 0.3, 0.9, 0.99):
            rho = 2 * eta * frac
            F, G, Fp, Gp = fs.coulomb_fg_below_turning_point(9, eta, rho)
            for l in (0, 1, 4, 9):
                f, g = mp.coulombf(l, eta, rho), mp.coulombg(l, eta, rho)
                worst = max(worst, float(abs(F[l] - f) / abs(f)), float(abs(G[l] - g) / abs(g)))
                assert abs(F[l] * Gp[l] - Fp[l] * G[l] + 1.0) < 1e-10  # Wronskian F G' - F' G = -1
    assert worst < 1e-13, worst

    def boom(*a, **k):
        raise AssertionError("mpmath reached")

    want = fs.coulomb_hankel_table(6, 1.413707698547276, 1.1)  # eta of examples/coupled.py, well inside the turning point
    ref = np.array([[complex(mp.coulombg(l, 1.413707698547276, 1.1)) + 1j * complex(mp.coulombf(l, 1.413707698547276, 1.1)) for l in range(7)]])
    assert np.max(np.abs(want[:, 0] - ref[0]) / np.abs(ref[0])) < 1e-13
    fs._TABLE_CACHE.clear()
    monkeypatch.setattr(fs.CoulombAsymptotics, "F", staticmethod(boom))
    monkeypatch.setattr(fs.CoulombAsymptotics, "G", staticmethod(boom))
    again = fs.coulomb_hankel_table(6, 1.413707698547276, 1.1)
    assert np.array_equal(again, want)
