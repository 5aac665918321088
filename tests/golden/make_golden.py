"""Generate golden vectors by running the REAL reference (jitr at /root/reference).

Run in the build container only (the reference cannot travel to the GPU box):

    python tests/golden/make_golden.py

Outputs small ``.npz`` fixtures in this directory.  They hold the exact inputs
(kinematics, masses, asymptotics, mesh potentials, parameter vectors) and the
reference's outputs (R, S, coefficients, cross sections) for each BASELINE
config and for restated versions of the reference's own unit tests.  The oracle
(``oracle/jitr_oracle.py``) is pinned to these; the CUDA path is compared with
both.  Versions used: see ``versions`` inside ``meta.npz``.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_env  # noqa: E402

_ref_env.setup()

import numpy as np  # noqa: E402

import jitr  # noqa: E402
from jitr import rmatrix  # noqa: E402
from jitr.optical_potentials import chuq, kduq, omp  # noqa: E402
from jitr.optical_potentials import potential_forms as pf  # noqa: E402
from jitr.reactions import ElasticReaction, ProjectileTargetSystem  # noqa: E402
from jitr.utils import kinematics as kin_mod  # noqa: E402
from jitr.utils.utils import interaction_range  # noqa: E402
from jitr.xs.elastic import DifferentialWorkspace  # noqa: E402


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")


def asym_table(ws):
    iw = ws.integral_workspace
    out = np.zeros((iw.lmax + 1, 4), dtype=np.complex128)
    for l in range(iw.lmax + 1):
        a = iw.asymptotics[l][0]
        out[l] = [a.Hp[0], a.Hm[0], a.Hpp[0], a.Hmp[0]]
    return out


def kin_array(k):
    return np.array([k.Elab, k.Ecm, k.mu, k.k, k.eta], dtype=np.float64)


def elastic_case(target, projectile, Elab, N, lmax, Rch, angles, model, params_list):
    """One workspace; many parameter vectors."""
    rx = ElasticReaction(target=target, projectile=projectile)
    kn = rx.kinematics(Elab)
    solver = rmatrix.Solver(N)
    ws = DifferentialWorkspace.build_from_system(rx, kn, Rch, solver, lmax, angles)
    rgrid = ws.radial_grid()
    outs = dict(dsdo=[], Ay=[], Q=[], tot=[], rxn=[], nl=[], central=[], so=[], coul=[])
    splus = np.zeros((len(params_list), lmax + 1), dtype=np.complex128)
    sminus = np.zeros_like(splus)
    for i, p in enumerate(params_list):
        c, so, co = model(rgrid, rx, kn, *p)
        sp, sm = ws.integral_workspace.smatrix(c, so, co)
        x = ws.xs(c, so, co)
        splus[i, : len(sp)] = sp
        sminus[i, : len(sm)] = sm
        outs["nl"].append(len(sp))
        outs["dsdo"].append(x.dsdo)
        outs["Ay"].append(x.Ay)
        outs["Q"].append(x.Q)
        outs["tot"].append(x.t)
        outs["rxn"].append(x.rxn)
        outs["central"].append(c)
        outs["so"].append(so)
        outs["coul"].append(np.asarray(co, dtype=np.float64) * np.ones(N))
    d = {k: np.array(v) for k, v in outs.items()}
    d.update(
        splus=splus, sminus=sminus, kin=kin_array(kn), masses=np.array([rx.target.m0, rx.projectile.m0]),
        target=np.array(target), projectile=np.array(projectile), N=N, lmax=lmax, Rch=Rch, angles=angles,
        asym=asym_table(ws), rgrid=rgrid, params=np.array(params_list),
        P=ws.P_l_costheta, P1=ws.P_1_l_costheta, sigma_l=ws.sigma_l, f_c=np.asarray(ws.f_c, dtype=np.complex128),
        rutherford=np.zeros(len(angles)) if ws.rutherford is None else ws.rutherford,
        free0=ws.integral_workspace.free_matrices[0], free3=ws.integral_workspace.free_matrices[3],
        basis_boundary=ws.integral_workspace.basis_boundary,
    )
    return d


def main():
    angles = np.linspace(0.1, np.pi, 180)
    versions = dict(
        numpy=np.__version__, scipy=__import__("scipy").__version__,
        numba=__import__("numba").__version__, mpmath=__import__("mpmath").__version__,
    )
    save("meta", versions=np.array(str(versions)))

    # ---- posterior samples / default parameter vectors (inputs of the sweeps) ----
    kn_samples = kduq.get_samples((1, 0))
    kp_samples = kduq.get_samples((1, 1))
    save(
        "kduq_params",
        kd03_n=kduq.get_kd03((1, 0)), kd03_p=kduq.get_kd03((1, 1)),
        federal_n=kn_samples, federal_p=kp_samples,
        names_n=np.array(kduq.get_param_names((1, 0))), names_p=np.array(kduq.get_param_names((1, 1))),
    )
    ch_samples = chuq.get_samples()
    save("chuq_params", ch89=chuq.get_ch89(), federal=ch_samples, names=np.array(chuq.get_param_names()))

    # ---- config 1: n+208Pb 30 MeV, N=40, lmax=30, KD03 (+ first 24 posterior rows) ----
    plist = [kduq.get_kd03((1, 0))] + [kn_samples[i] for i in range(24)]
    save("config1_n208Pb_30MeV", **elastic_case((208, 82), (1, 0), 30.0, 40, 30, 12.0, angles, kduq.KDUQ((1, 0)), plist))

    # ---- config 3: p+48Ca 30 MeV, N=60, lmax=40, KD03 proton (+ 16 posterior rows) ----
    plist = [kduq.get_kd03((1, 1))] + [kp_samples[i] for i in range(16)]
    save("config3_p48Ca_30MeV", **elastic_case((48, 20), (1, 1), 30.0, 60, 40, 12.0, angles, kduq.KDUQ((1, 1)), plist))

    # ---- config 2 subset: n+40Ca / n+208Pb at 4 of the 32 energies, suggested radius ----
    egrid = np.linspace(5.0, 67.0, 32)
    for tgt, tag in (((40, 20), "n40Ca"), ((208, 82), "n208Pb")):
        for ie in (0, 6, 15, 31):
            Elab = float(egrid[ie])
            rx = ElasticReaction(target=tgt, projectile=(1, 0))
            kn = rx.kinematics(Elab)
            Rch = (interaction_range(tgt[0]) * kn.k + 2 * np.pi) / kn.k
            plist = [kn_samples[i] for i in range(0, 416, 32)]
            save(f"config2_{tag}_E{ie:02d}", ie=ie, **elastic_case(tgt, (1, 0), Elab, 40, 30, Rch, angles, kduq.KDUQ((1, 0)), plist))

    # ---- CHUQ: p+90Zr 25 MeV and n+90Zr 12 MeV, N=40, lmax=30 ----
    plist = [chuq.get_ch89()] + [ch_samples[i] for i in range(8)]

    class _CH(omp.SingleChannelOpticalModel):
        def __init__(self):
            super().__init__(chuq.get_param_names())

        def evaluate(self, rgrid, reaction, kinematics, *p):
            c, s, co = chuq.calculate_params(
                (reaction.projectile.A, reaction.projectile.Z), (reaction.target.A, reaction.target.Z),
                kinematics.Elab, *p)
            return chuq.central(rgrid, *c), chuq.spin_orbit(rgrid, *s), pf.coulomb_charged_sphere(rgrid, *co)

    save("chuq_p90Zr_25MeV", **elastic_case((90, 40), (1, 1), 25.0, 40, 30, 14.0, angles, _CH(), plist))
    save("chuq_n90Zr_12MeV", **elastic_case((90, 40), (1, 0), 12.0, 40, 30, 14.0, angles, _CH(), plist))

    # ---- LocalOpticalPotential: p+54Fe 35 MeV ----
    lp = [
        np.array([50.0, 1.2, 0.65, 3.0, 1.25, 0.6, 6.0, 1.5, 1.3, 0.55, 5.5, -0.2, 1.1, 0.6, 1.3]),
        np.array([46.0, 1.17, 0.7, 1.0, 1.3, 0.55, 8.0, 0.0, 1.28, 0.5, 6.0, 0.0, 1.0, 0.65, 1.25]),
    ]
    save("localomp_p54Fe_35MeV", **elastic_case((54, 26), (1, 1), 35.0, 40, 50, 13.0, angles[::3], omp.LocalOpticalPotential(), lp))

    # ---- KDUQ shape parameters for all 416 rows at several (target, Elab) ----
    shp = {}
    for proj, samples, tag in (((1, 0), kn_samples, "n"), ((1, 1), kp_samples, "p")):
        for tgt in ((40, 20), (208, 82), (48, 20)):
            for Elab in (5.0, 30.0, 67.0):
                rows = []
                for p in samples:
                    c, s, co = kduq.calculate_params(proj, tgt, Elab, *p)
                    rows.append(list(c) + list(s) + list(co))
                shp[f"{tag}_{tgt[0]}_{tgt[1]}_{Elab:g}"] = np.array(rows, dtype=np.float64)
    save("kduq_shape_params", **shp)

    # ---- restated tests/test_local.py: p+"Ca48" WS+Coulomb, N=70, a=10 pi, l=0..2, 10 energies ----
    mass_t, mass_p = 44657.26581995028, 938.271653086152
    sysl = ProjectileTargetSystem(channel_radius=10 * np.pi, lmax=2, mass_target=mass_t, mass_projectile=mass_p, Ztarget=20, Zproj=1)
    solver = rmatrix.Solver(70)
    egrid10 = np.linspace(0.5, 100, 10)
    prm = (60, 20, 4, 0.5, 20, 4)
    S = np.zeros((10, 3), dtype=np.complex128)
    R = np.zeros_like(S)
    U = np.zeros_like(S)
    kins = np.zeros((10, 5))
    asy = np.zeros((10, 3, 4), dtype=np.complex128)
    pots = np.zeros((10, 70), dtype=np.complex128)
    for i, Elab in enumerate(egrid10):
        kk = kin_mod.classical_kinematics(mass_t, mass_p, Elab, 20)
        kins[i] = kin_array(kk)
        channels, asymptotics = sysl.get_partial_wave_channels(*kk)
        for l in range(3):
            rgrid = solver.radial_grid(channels[l].a, channels[l].k[0])
            V = -pf.woods_saxon_potential(rgrid, *prm[:4]) + pf.coulomb_charged_sphere(rgrid, prm[4], prm[5])
            pots[i] = V
            r, s, u = solver.solve(channels[l], asymptotics[l], local_potential=V)
            R[i, l], S[i, l], U[i, l] = r[0, 0], s[0, 0], u[0]
            a_ = asymptotics[l]
            asy[i, l] = [a_.Hp[0], a_.Hm[0], a_.Hpp[0], a_.Hmp[0]]
    save("test_local", kin=kins, asym=asy, pot=pots, R=R, S=S, uext=U, a=10 * np.pi, N=70, masses=np.array([mass_t, mass_p]))

    # ---- restated tests/test_wavefunction.py: N=100, l=0, wavefunction=True ----
    sysw = ProjectileTargetSystem(channel_radius=10 * np.pi, lmax=10, mass_target=mass_t, mass_projectile=mass_p, Ztarget=20, Zproj=1)
    kk = kin_mod.classical_kinematics(mass_t, mass_p, 14.1, 20)
    channels, asymptotics = sysw.get_partial_wave_channels(*kk)
    solver = rmatrix.Solver(100)
    wf = {}
    for l in (0, 1, 4):
        rgrid = solver.radial_grid(channels[l].a, channels[l].k[0])
        V = pf.coulomb_charged_sphere(rgrid, 20, 6) - pf.woods_saxon_potential(rgrid, 70, 40, 6, 1.2)
        r, s, x, u = solver.solve(channels[l], asymptotics[l], wavefunction=True, local_potential=V)
        a_ = asymptotics[l]
        wf[f"R{l}"], wf[f"S{l}"], wf[f"x{l}"], wf[f"u{l}"] = r, s, x, u
        wf[f"asym{l}"] = np.array([a_.Hp[0], a_.Hm[0], a_.Hpp[0], a_.Hmp[0]])
        wf[f"pot{l}"] = V
        from jitr.reactions import wavefunction as wfm

        svals = np.linspace(0.01, sysw.channel_radius, 50)
        wf[f"uint{l}"] = wfm.Wavefunctions(solver, x, s, u, channels[l]).uint()[0](svals)
        wf["svals"] = svals
    save("test_wavefunction", kin=kin_array(kk), a=10 * np.pi, N=100, **wf)

    # ---- config 5: coupled 3-channel (examples/coupled.py geometry), N=50 and N=40, l = 0..3 ----
    from jitr.utils import constants, mass

    levels = np.array([0, 2.3, 3.1])
    mass_Ca48 = mass.mass(48, 20)[0]
    cm = np.array([[0, 0.8, 0.1], [0, 0, 0.1], [0, 0, 0.0]])
    cm = cm + cm.T
    sysc = ProjectileTargetSystem(
        channel_radius=6 * np.pi, lmax=10, mass_target=mass_Ca48, mass_projectile=constants.MASS_P,
        Ztarget=20, Zproj=1, channel_levels=levels, coupling=lambda l: cm,
    )
    kk = kin_mod.classical_kinematics(sysc.mass_target, sysc.mass_projectile, 5, 20)
    kk.Ecm -= sysc.channel_levels
    channels, asymptotics = sysc.get_partial_wave_channels(*kk)
    cc = dict(E=channels[0].E, k=channels[0].k, mu=channels[0].mu, eta=channels[0].eta, a=channels[0].a, coupling=cm)
    for N in (40, 50):
        solver = rmatrix.Solver(N)
        for l in range(4):
            rgrid = solver.radial_grid(channels[l].a, channels[l].k[0])
            coul = pf.coulomb_charged_sphere(rgrid, 1, 4)
            nuc = pf.woods_saxon_potential(rgrid, -42, 0, 4, 0.8)
            z = np.zeros_like(rgrid)
            diag = np.array([[nuc + coul, z, z], [z, nuc + coul, z], [z, z, nuc + coul]])
            V = diag + cm[..., None] * pf.surface_peaked_gaussian_potential(rgrid, -42, 0, 4, 0.8)
            r, s, x, u = solver.solve(channels[l], asymptotics[l], local_potential=V, wavefunction=True)
            a_ = asymptotics[l]
            cc[f"N{N}_l{l}_R"], cc[f"N{N}_l{l}_S"], cc[f"N{N}_l{l}_x"], cc[f"N{N}_l{l}_u"] = r, s, x, u
            cc[f"N{N}_l{l}_V"] = V
            cc[f"l{l}_asym"] = np.array([a_.Hp, a_.Hm, a_.Hpp, a_.Hmp])
    save("config5_coupled", **cc)

    # ---- restated tests/test_coupled_single_dwba.py: 2 uncoupled channels, N=40 ----
    sys2 = ProjectileTargetSystem(
        channel_radius=5 * np.pi, lmax=10, mass_target=44657, mass_projectile=938.3, Ztarget=20, Zproj=1,
        coupling=lambda l: np.array([[1, 0], [0, 1]]),
    )
    kk = kin_mod.classical_kinematics(44657, 938.3, 42.1, 20)
    channels, asymptotics = sys2.get_partial_wave_channels(*kk)
    solver = rmatrix.Solver(40)
    rgrid = solver.radial_grid(channels[0].a, channels[0].k[0])
    d = -10 * np.exp(-rgrid / 4)
    o = -0 * np.exp(-rgrid / 4)
    V2 = np.array([[d, o], [o, d]])
    Rm, Sm, um = solver.solve(channels[0], asymptotics[0], local_potential=V2)
    ch0 = channels[0].decouple()[0]
    as0 = asymptotics[0].decouple()[0]
    R1, S1, u1 = solver.solve(ch0, as0, local_potential=d)
    a_ = asymptotics[0]
    save("test_coupled_single", kin=kin_array(kk), a=5 * np.pi, N=40, V2=V2, V1=d, Rm=Rm, Sm=Sm, um=um, R1=R1, S1=S1, u1=u1,
         asym=np.array([a_.Hp, a_.Hm, a_.Hpp, a_.Hmp]))

    # ---- nonlocal: Yamaguchi (examples/nonlocal.py) N=20,40; Perey-Buck-style N=60 kernel (config 4) ----
    W0, beta, alpha = 41.472, 1.3918324, 0.2316053
    ecom = 12
    mu = constants.HBARC**2 / (2 * W0)
    k = np.sqrt(2 * mu * ecom) / constants.HBARC
    sysn = ProjectileTargetSystem(channel_radius=30.0, lmax=5)
    channels, asymptotics = sysn.get_partial_wave_channels(ecom, ecom, mu, k, 0)
    nl = dict(kin=np.array([ecom, ecom, mu, k, 0.0]), a=30.0)
    for N in (20, 40):
        solver = rmatrix.Solver(N)
        rg, rpg = solver.nonlocal_radial_grids(channels[0].a, channels[0].k[0])
        V = pf.yamaguchi_potential(rg, rpg, W0, beta, alpha)
        r, s, u = solver.solve(channels[0], asymptotics[0], nonlocal_potential=V)
        nl[f"yam{N}_V"], nl[f"yam{N}_R"], nl[f"yam{N}_S"], nl[f"yam{N}_u"] = V, r, s, u
    a_ = asymptotics[0]
    nl["yam_asym"] = np.array([a_.Hp[0], a_.Hm[0], a_.Hpp[0], a_.Hmp[0]])
    save("nonlocal_yamaguchi", **nl)

    # config 4: n+208Pb 30 MeV, N=60; U(r,r') = -WS((r+r')/2) H_l(r,r';beta) built stably with ive
    from scipy.special import ive

    rx = ElasticReaction(target=(208, 82), projectile=(1, 0))
    kn = rx.kinematics(30.0)
    N, lmax, Rch = 60, 6, 12.0
    a = Rch * kn.k
    sys4 = ProjectileTargetSystem(a, lmax, mass_target=rx.target.m0, mass_projectile=rx.projectile.m0, Ztarget=82, Zproj=0)
    channels, asymptotics = sys4.get_partial_wave_channels(*kn)
    solver = rmatrix.Solver(N)
    rg, rpg = solver.nonlocal_radial_grids(a, kn.k)
    c4 = dict(kin=kin_array(kn), a=a, N=N, Rch=Rch)
    for isamp, (Vd, Wd, bt) in enumerate([(71.0, 15.0, 0.85), (68.2, 13.1, 0.88)]):
        for l in range(lmax + 1):
            z = 2 * rg * rpg / bt**2
            # 2 i^l z j_l(-i z) = 2 z i_l(z) (modified spherical Bessel) ; i_l(z) = sqrt(pi/(2z)) I_{l+1/2}(z)
            # exp(-(r^2+r'^2)/b^2) * 2 z i_l(z) = 2 z sqrt(pi/(2z)) ive(l+1/2, z) exp(-(r-r')^2/b^2)
            H = 2 * z * np.sqrt(np.pi / (2 * z)) * ive(l + 0.5, z) * np.exp(-((rg - rpg) ** 2) / bt**2) / (bt * np.sqrt(np.pi))
            U = -pf.woods_saxon_potential((rg + rpg) / 2, Vd, Wd, 1.25 * 208 ** (1 / 3), 0.65) * H
            r, s, u = solver.solve(channels[l], asymptotics[l], nonlocal_potential=U)
            c4[f"s{isamp}_l{l}_S"], c4[f"s{isamp}_l{l}_R"] = s, r
            if l in (0, 3):
                c4[f"s{isamp}_l{l}_U"] = U
            a_ = asymptotics[l]
            c4[f"l{l}_asym"] = np.array([a_.Hp[0], a_.Hm[0], a_.Hpp[0], a_.Hmp[0]])
    c4["samples"] = np.array([(71.0, 15.0, 0.85), (68.2, 13.1, 0.88)])
    save("config4_nonlocal", **c4)

    # ---- notebook golden: kduq_uq_demo.ipynb cell 20 (84th percentile of dsdo/ruth at angle index 9) ----
    # p+54Fe 35 MeV, N=40, lmax=50, 180 angles, all 416 proton federal samples
    rx = ElasticReaction(target=(54, 26), projectile=(1, 1))
    kn = rx.kinematics(35.0)
    Rch = (interaction_range(54) * kn.k + 2 * np.pi) / kn.k
    ang = np.linspace(0.1, np.pi, 180)
    solver = rmatrix.Solver(40)
    ws = DifferentialWorkspace.build_from_system(rx, kn, Rch, solver, 50, ang)
    model = kduq.KDUQ((1, 1))
    rat = np.zeros((416, 180))
    for i, p in enumerate(kp_samples):
        x = ws.xs(*model(ws.radial_grid(), rx, kn, *p))
        rat[i] = x.dsdo / ws.rutherford
    save("nb_kduq_uq_demo", ratio=rat, kin=kin_array(kn), Rch=Rch, masses=np.array([rx.target.m0, rx.projectile.m0]),
         asym=asym_table(ws), pct84_idx9=np.percentile(rat, 84, axis=0)[9])
    print("kduq_uq_demo 84th pct @ idx 9:", np.percentile(rat, 84, axis=0)[9])


if __name__ == "__main__":
    main()
