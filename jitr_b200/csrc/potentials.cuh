// Built-in optical-potential forms and global-parameter maps, evaluated on the mesh inside the
// fused solve kernel.  Reference: optical_potentials/potential_forms.py:11,40-88,134-153 (forms),
// kduq.py:100-131,370-528 (KDUQ map), chuq.py:140-227 (CHUQ map), wlh.py:251-383 (WLH map),
// omp.py:54-149 (Local).
#pragma once
#include <cuda_runtime.h>
#include "jitr_b200.h"

namespace jitr {

#define ET_A13 JITR_B200_ET_A13
#define ET_AM13 JITR_B200_ET_AM13
#define ET_AM23 JITR_B200_ET_AM23
#define ET_AM53 JITR_B200_ET_AM53

constexpr double kHbarc = 197.3269804;          // utils/constants.py:5
constexpr double kAlpha = 1.0 / 137.0359991;    // utils/constants.py:6
constexpr double kKappa2 = 0.5000000000000001;  // WAVENUMBER_PION**2 = sqrt(1/2)**2 (constants.py:7)
constexpr double kMaxArg = 36.841361487904734;  // log(1/1e-16), potential_forms.py:11

// canonical shape vector:
//  V_c  = -Vv f(Rv,av) - (Vw + i Wv) f(Rw,aw) + 4 ad (Vd + i Wd) f'(Rd,ad)     (Vw: dispersive correction of the
//         real depth at the imaginary-volume geometry, dom.py:100-104; 0 for the other models)
//  V_so = Vso g(Rso,aso) + i Wso g(Rwso,awso)        (Vso, Wso already include 1/kappa^2 or 2)
//  V_C  = Zz alpha hbarc * (r <= RC ? (3 - (r/RC)^2)/(2 RC) : 1/r)
struct Shape {
    double Vv, Rv, av, Wv, Rw, aw, Vd, Wd, Rd, ad, Vso, Rso, aso, Wso, Rwso, awso, RC, Zz;
    double Vw = 0.0;
};

__device__ __forceinline__ void ws_forms(double r, double R, double a, double& f, double& fp) {
    // woods_saxon_safe / woods_saxon_prime_safe (potential_forms.py:40-67): exactly 0 beyond MAX_ARG
    double x = (r - R) / a;
    if (x <= kMaxArg) {
        double e = exp(x);
        double d = 1.0 + e;
        f = 1.0 / d;
        fp = -e / (a * (d * d));
    } else {
        f = 0.0;
        fp = 0.0;
    }
}

__device__ __forceinline__ void eval_shape(const Shape& s, double r, double2& vc, double2& vso, double& vcoul) {
    // The built-in models share geometries between terms (KDUQ: volume real/imaginary and both
    // spin-orbit terms; CHUQ: imaginary volume/surface; Local: both spin-orbit terms): each distinct
    // (R, a) is evaluated once.  The comparisons are warp-uniform.
    double fv, fpv, fw, fpw, fd, fpd, fs, fps, fws, fpws;
    ws_forms(r, s.Rv, s.av, fv, fpv);
    if (s.Rw == s.Rv && s.aw == s.av) { fw = fv; fpw = fpv; } else ws_forms(r, s.Rw, s.aw, fw, fpw);
    if (s.Rd == s.Rw && s.ad == s.aw) { fd = fw; fpd = fpw; } else ws_forms(r, s.Rd, s.ad, fd, fpd);
    ws_forms(r, s.Rso, s.aso, fs, fps);
    if (s.Rwso == s.Rso && s.awso == s.aso) { fws = fs; fpws = fps; } else ws_forms(r, s.Rwso, s.awso, fws, fpws);
    (void)fd; (void)fs; (void)fws;
    double re = -s.Vv * fv;
    if (s.Vw != 0.0) re -= s.Vw * fw;  // warp-uniform
    double im = -s.Wv * fw;
    // - (-4 ad) Vd f'  - i (-4 ad) Wd f'   (omp.py:68-73, kduq.py:163-167)
    re -= (-4.0 * s.ad) * s.Vd * fpd;
    im -= (-4.0 * s.ad) * s.Wd * fpd;
    vc = make_double2(re, im);
    // thomas_safe = (1/r) f'  (potential_forms.py:70-88)
    double y = 1.0 / r;
    vso = make_double2(s.Vso * (y * fps), s.Wso * (y * fpws));
    if (s.Zz == 0.0) {
        vcoul = 0.0;
    } else {
        double inv = (r <= s.RC) ? 1.0 / (2.0 * s.RC) * (3.0 - (r / s.RC) * (r / s.RC)) : 1.0 / r;
        vcoul = s.Zz * kAlpha * kHbarc * inv;
    }
}

// ---- global parameter maps -------------------------------------------------------------------
__device__ __forceinline__ Shape shape_from_canonical(const double* __restrict__ p, const double* __restrict__ et) {
    Shape s;
    s.Vv = p[0]; s.Rv = p[1]; s.av = p[2]; s.Wv = p[3]; s.Rw = p[4]; s.aw = p[5];
    s.Vd = p[6]; s.Wd = p[7]; s.Rd = p[8]; s.ad = p[9];
    s.Vso = p[10]; s.Rso = p[11]; s.aso = p[12]; s.Wso = p[13]; s.Rwso = p[14]; s.awso = p[15];
    s.RC = p[16];
    s.Zz = et[JITR_B200_ET_ZT] * et[JITR_B200_ET_ZP];
    return s;
}

// kduq.calculate_params, kduq.py:438-528 ; p has 40 entries (neutron rows padded with zeros)
__device__ __forceinline__ Shape shape_from_kduq(const double* __restrict__ p, const double* __restrict__ et) {
    const double A = et[JITR_B200_ET_AT], Z = et[JITR_B200_ET_ZT], Zp = et[JITR_B200_ET_ZP];
    const double Elab = et[JITR_B200_ET_ELAB];
    const double A13 = et[ET_A13], Am13 = et[ET_AM13], Am23 = et[ET_AM23], Am53 = et[ET_AM53];
    double asym = (A - 2.0 * Z) / A;
    const double factor = (Zp == 1.0) ? 1.0 : -1.0;  // (-1)**(Zp+1)
    asym *= factor;
    const double Ef = p[0] + p[1] * A;
    const double v1 = p[2] + p[3] * asym - p[4] * A;
    const double v2 = p[5] + p[6] * A * factor;
    const double v3 = p[7] + p[8] * A * factor;
    const double v4 = p[9];
    const double dE = Elab - Ef;
    double vv = v1 * (1.0 - v2 * dE + v3 * (dE * dE) - v4 * (dE * dE * dE));
    const double rv = p[10] - p[11] * Am13;
    const double av = p[12] - p[13] * A;
    const double w1 = p[14] + p[15] * A;
    const double w2 = p[16] + p[17] * A;
    const double wv = w1 * (dE * dE) / (dE * dE + w2 * w2);
    const double d1 = p[18] + p[19] * asym;
    const double d2 = p[20] + p[21] / (1.0 + exp((A - p[23]) / p[22]));  // inf -> 0 like numpy (kduq.py:470)
    const double d3 = p[24];
    const double wd = d1 * (dE * dE) / (dE * dE + d3 * d3) * exp(-d2 * dE);
    const double rd = p[25] - p[26] * A13;
    const double ad = p[27] + p[28] * A * factor;
    const double vso1 = p[29] + p[30] * A;
    const double vso = vso1 * exp(-p[31] * dE);
    const double wso = p[32] * (dE * dE) / (dE * dE + p[33] * p[33]);
    const double rso = p[34] - p[35] * Am13;
    const double aso = p[36];
    double RC = rv * A13;
    if (Zp == 1.0) {
        const double rc0 = p[37] + p[38] * Am23 + p[39] * Am53;
        RC = rc0 * A13;
        const double Vcbar = 1.73 / rc0 * Z * Am13;
        vv += v1 * Vcbar * (v2 - 2.0 * v3 * dE + 3.0 * v4 * (dE * dE));
    }
    Shape s;
    s.Vv = vv; s.Rv = rv * A13; s.av = av;
    s.Wv = wv; s.Rw = rv * A13; s.aw = av;
    s.Vd = 0.0; s.Wd = wd; s.Rd = rd * A13; s.ad = ad;
    s.Vso = vso / kKappa2; s.Rso = rso * A13; s.aso = aso;
    s.Wso = wso / kKappa2; s.Rwso = rso * A13; s.awso = aso;
    s.RC = RC;
    s.Zz = Z * Zp;
    return s;
}

// chuq.calculate_params, chuq.py:140-227 ; 22 entries
__device__ __forceinline__ Shape shape_from_chuq(const double* __restrict__ p, const double* __restrict__ et) {
    const double A = et[JITR_B200_ET_AT], Z = et[JITR_B200_ET_ZT], Zp = et[JITR_B200_ET_ZP];
    const double Elab = et[JITR_B200_ET_ELAB];
    const double A13 = et[ET_A13];
    const bool is_p = (Zp == 1.0);
    const double alpha = ((A - Z) - Z) / A;
    const double asym = (is_p ? 1.0 : -1.0) * alpha;
    const double R0 = p[3] * A13 + p[4];
    const double Rw = p[9] * A13 + p[10];
    const double Rso = p[17] * A13 + p[18];
    const double RC = p[20] * A13 + p[21];
    const double Ec = is_p ? 6.0 * Z * kAlpha * kHbarc / (5.0 * RC) : 0.0;
    const double dE = Elab - Ec;
    const double V = p[0] + p[1] * dE + asym * p[2];
    const double Wv = p[6] / (1.0 + exp((p[7] - dE) / p[8]));
    const double Ws = (p[12] + asym * p[13]) / (1.0 + exp((dE - p[14]) / p[15]));
    Shape s;
    s.Vv = V; s.Rv = R0; s.av = p[5];
    s.Wv = Wv; s.Rw = Rw; s.aw = p[11];
    s.Vd = 0.0; s.Wd = Ws; s.Rd = Rw; s.ad = p[11];
    s.Vso = 2.0 * p[16]; s.Rso = Rso; s.aso = p[19];
    s.Wso = 0.0; s.Rwso = Rso; s.awso = p[19];
    s.RC = RC;
    s.Zz = Z * Zp;
    return s;
}

// LocalOpticalPotential.evaluate, omp.py:115-149 ; 15 entries, radius factor from the energy table
__device__ __forceinline__ Shape shape_from_local(const double* __restrict__ p, const double* __restrict__ et) {
    const double rf = et[JITR_B200_ET_RSCALE];
    Shape s;
    s.Vv = p[0]; s.Rv = p[1] * rf; s.av = p[2];
    s.Wv = p[3]; s.Rw = p[4] * rf; s.aw = p[5];
    s.Wd = p[6]; s.Vd = p[7]; s.Rd = p[8] * rf; s.ad = p[9];
    // (Vso + i Wso)/kappa^2 * thomas(Rso, aso)  (omp.py:79-81)
    s.Vso = p[10] / kKappa2; s.Rso = p[12] * rf; s.aso = p[13];
    s.Wso = p[11] / kKappa2; s.Rwso = p[12] * rf; s.awso = p[13];
    s.RC = p[14] * rf;
    s.Zz = et[JITR_B200_ET_ZT] * et[JITR_B200_ET_ZP];
    return s;
}

// wlh.calculate_params, wlh.py:251-383 ; 46 entries (get_param_names order)
__device__ __forceinline__ Shape shape_from_wlh(const double* __restrict__ p, const double* __restrict__ et) {
    const double A = et[JITR_B200_ET_AT], Z = et[JITR_B200_ET_ZT], Zp = et[JITR_B200_ET_ZP];
    const double E = et[JITR_B200_ET_ELAB];
    const double A13 = et[ET_A13], Am13 = et[ET_AM13];
    const double asym = (A - 2.0 * Z) / A;
    const double factor = (Zp == 1.0) ? 1.0 : -1.0;  // (-1)**(Zp+1)
    const double E2 = E * E, E3 = E * E * E;
    const double uv = p[0] - p[1] * E + p[2] * E2 + p[3] * E3 + factor * (p[4] - p[5] * E + p[6] * E2) * asym;
    const double rv = p[7] - p[8] * Am13 - p[9] * E + p[10] * E2;
    const double av = p[11] - factor * p[12] * E - p[13] * E2 - (p[14] - p[15] * asym) * asym;
    const double uw = p[16] + p[17] * E - p[18] * E2 + (factor * p[19] - p[20] * E) * asym;
    const double rw = p[21] + (p[22] + p[23] * A) / (p[24] + A + p[25] * E) + p[26] * E2;
    const double aw = p[27] - (p[28] * E) / (-p[29] - E) + (p[30] - p[31] * E) * asym;
    double ud = 0.0;
    if ((Zp == 0.0 && E < 40.0) || (Zp == 1.0 && E < 20.0 && A > 100.0))
        ud = -(p[32] - p[33] * E - (p[34] - p[35] * E) * asym);
    const double rd = p[36] - p[38] * E - p[37] * Am13;
    const double ad = p[39];
    const double uso = p[40] - p[41] * A;
    const double rso = p[42] - p[43] * Am13;
    const double aso = p[44] - p[45] * A;
    Shape s;
    s.Vv = uv; s.Rv = rv * A13; s.av = av;
    s.Wv = uw; s.Rw = rw * A13; s.aw = aw;
    s.Vd = 0.0; s.Wd = ud; s.Rd = rd * A13; s.ad = ad;
    s.Vso = uso / kKappa2; s.Rso = rso * A13; s.aso = aso;
    s.Wso = 0.0; s.Rwso = rso * A13; s.awso = aso;
    s.RC = rv * A13;
    s.Zz = Z * Zp;
    return s;
}

// 17 canonical numbers + Vw, one row per (sample, energy): models whose parameter map stays on the host (DOM)
__device__ __forceinline__ Shape shape_from_canonical_pair(const double* __restrict__ p, const double* __restrict__ et) {
    Shape s = shape_from_canonical(p, et);
    s.Vw = p[17];
    return s;
}

// ---- dispersive optical model (optical_potentials/dom.py:151-283, 287-382, 399-455) --------------------------------
// The energy dependence of the DOM depths contains the analytic surface-dispersion integral: exponential integrals of real
// argument, and of complex argument on the rays arg z = -pi/4, -3pi/4.  Series below |z| <= 8 (x <= 1 for the real E1, whose
// alternating series cancels beyond), continued fractions (modified Lentz) / the asymptotic series above; all in the scaled
// forms e^{-x} Ei(x), e^{x} E1(x), e^{z} E1(z) that the formula uses, so nothing overflows.
constexpr double kEulerGamma = 0.5772156649015329;

static __device__ __noinline__ double dom_expi_scaled(double x) {  // e^{-x} Ei(x), x > 0
    if (x <= 40.0) {
        double term = 1.0, sum = 0.0;
        for (int k = 1; k < 200; ++k) {
            term *= x / k;
            const double t = term / k;
            sum += t;
            if (t < 1e-17 * sum) break;
        }
        return exp(-x) * (kEulerGamma + log(x) + sum);
    }
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 60; ++k) {
        const double next = term * k / x;
        if (next >= term || next < 1e-17) break;
        term = next;
        sum += term;
    }
    return sum / x;
}

static __device__ __noinline__ double2 dom_cexp1_scaled(double2 z) {  // e^{z} E1(z), z off the negative real axis
    const double r2 = z.x * z.x + z.y * z.y;
    if (r2 <= 64.0 && !(z.y == 0.0 && z.x > 1.0)) {
        // E1(z) = -gamma - Log z - sum_{k>=1} (-z)^k / (k k!)
        double2 term = make_double2(1.0, 0.0), sum = make_double2(0.0, 0.0);
        for (int k = 1; k < 200; ++k) {
            const double tx = -(term.x * z.x - term.y * z.y) / k, ty = -(term.x * z.y + term.y * z.x) / k;
            term = make_double2(tx, ty);
            sum.x += tx / k;
            sum.y += ty / k;
            if (fabs(tx) + fabs(ty) < 1e-17 * (fabs(sum.x) + fabs(sum.y)) * k) break;
        }
        const double2 e1 = make_double2(-kEulerGamma - 0.5 * log(r2) - sum.x, -atan2(z.y, z.x) - sum.y);
        const double ex = exp(z.x);
        const double2 ez = make_double2(ex * cos(z.y), ex * sin(z.y));
        return make_double2(ez.x * e1.x - ez.y * e1.y, ez.x * e1.y + ez.y * e1.x);
    }
    // e^z E1(z) = 1 / (z + 1 - 1 / (z + 3 - 4 / (z + 5 - ...)))   (modified Lentz)
    const double tiny = 1e-300;
    double2 b = make_double2(z.x + 1.0, z.y);
    double2 f = b, C = b, D = make_double2(0.0, 0.0);
    auto cinv = [](double2 v) { const double d = v.x * v.x + v.y * v.y; return make_double2(v.x / d, -v.y / d); };
    for (int i = 1; i < 500; ++i) {
        const double an = -(double)i * (double)i;
        b.x += 2.0;
        D = make_double2(b.x + an * D.x, b.y + an * D.y);
        if (D.x == 0.0 && D.y == 0.0) D.x = tiny;
        const double2 Ci = cinv(C);
        C = make_double2(b.x + an * Ci.x, b.y + an * Ci.y);
        if (C.x == 0.0 && C.y == 0.0) C.x = tiny;
        D = cinv(D);
        const double2 delta = make_double2(C.x * D.x - C.y * D.y, C.x * D.y + C.y * D.x);
        f = make_double2(f.x * delta.x - f.y * delta.y, f.x * delta.y + f.y * delta.x);
        if (fabs(delta.x - 1.0) + fabs(delta.y) < 1e-16) break;
    }
    return cinv(f);
}

// delta_Vs_analytic, dom.py:399-455
static __device__ __noinline__ double dom_delta_Vs(double x, double w0, double s1, double s2) {
    if (x == 0.0) return 0.0;
    const double ax = fabs(x);
    double J1 = 0.0, J2;
    if (s2 != 0.0) {
        const double z1 = s2 * ax;
        J1 = (-dom_expi_scaled(z1) - dom_cexp1_scaled(make_double2(z1, 0.0)).x) / (2.0 * ax);
        J2 = 0.0;
        const double c = 0.7071067811865476;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            // root = s1 e^{i pi/4}, s1 e^{3 i pi/4};  F = e^{z} E1(z), z = -s2 root
            const double2 root = make_double2(m == 0 ? s1 * c : -s1 * c, s1 * c);
            const double2 F = dom_cexp1_scaled(make_double2(-s2 * root.x, -s2 * root.y));
            // (root^2 + x^2) / (4 root^3) * F
            const double2 r2 = make_double2(root.x * root.x - root.y * root.y, 2.0 * root.x * root.y);
            const double2 r3 = make_double2(r2.x * root.x - r2.y * root.y, r2.x * root.y + r2.y * root.x);
            const double2 num = make_double2(r2.x + x * x, r2.y);
            const double d = 4.0 * (r3.x * r3.x + r3.y * r3.y);
            const double2 q = make_double2((num.x * r3.x + num.y * r3.y) / d, (num.y * r3.x - num.x * r3.y) / d);
            J2 += 2.0 * (q.x * F.x - q.y * F.y);
        }
    } else {
        J2 = 3.141592653589793 * (s1 * s1 + x * x) / (2.0 * 1.4142135623730951 * s1 * s1 * s1);
    }
    const double x4 = x * x * x * x, s4 = s1 * s1 * s1 * s1;
    return (2.0 * w0 * x / 3.141592653589793) * (x4 / (x4 + s4) * J1 + s4 / (x4 + s4) * J2);
}

// dom.calculate_params, dom.py:287-382; p: the 19 global parameters of get_param_names(); the effective Fermi energy of the
// (projectile, target) pair comes with the energy row (JITR_B200_ET_EF)
static __device__ __noinline__ Shape shape_from_dom(const double* __restrict__ p, const double* __restrict__ et) {
    const double Z = et[JITR_B200_ET_ZT], Zp = et[JITR_B200_ET_ZP], A13 = et[ET_A13];
    const double v0 = p[0], v1 = p[1], rv = p[2], av = p[3], wv0 = p[4], wv1 = p[5], rw = p[6], aw = p[7], ws0 = p[8], ws1 = p[9],
                 ws2 = p[10], rd = p[11], ad = p[12], vso0 = p[13], wso0 = p[14], wso1 = p[15], rso = p[16], aso = p[17], rC = p[18];
    const double RC = rC * A13;
    const double Ec = (Zp == 1.0) ? 6.0 * (Z * Zp) * kAlpha * kHbarc / (5.0 * RC) : 0.0;  // dom.py:385-396
    const double dE = et[JITR_B200_ET_ECM] - Ec - et[JITR_B200_ET_EF];
    const double dE2 = dE * dE, dE4 = dE2 * dE2;
    Shape s;
    s.Vv = v0 * exp(-v1 * dE);
    s.Rv = rv * A13; s.av = av;
    s.Wv = wv0 * dE2 / (dE2 + wv1 * wv1);
    s.Rw = rw * A13; s.aw = aw;
    s.Vd = dom_delta_Vs(dE, ws0, ws1, ws2);
    s.Wd = ws0 * dE4 / (dE4 + ws1 * ws1 * ws1 * ws1) * exp(-ws2 * fabs(dE));
    s.Rd = rd * A13; s.ad = ad;
    s.Vso = 2.0 * (vso0 * exp(-v1 * dE) + wso0 * wso1 * dE / (dE2 + wso1 * wso1));
    s.Rso = rso * A13; s.aso = aso;
    s.Wso = 2.0 * (wso0 * dE2 / (dE2 + wso1 * wso1));
    s.Rwso = s.Rso; s.awso = aso;
    s.RC = RC;
    s.Zz = Z * Zp;
    s.Vw = wv0 * wv1 * dE / (dE2 + wv1 * wv1);
    return s;
}

__device__ __forceinline__ Shape shape_from_model(int model, const double* __restrict__ p, const double* __restrict__ et) {
    switch (model) {
        case JITR_B200_MODEL_SHAPE_PAIR: return shape_from_canonical_pair(p, et);
        case JITR_B200_MODEL_KDUQ: return shape_from_kduq(p, et);
        case JITR_B200_MODEL_CHUQ: return shape_from_chuq(p, et);
        case JITR_B200_MODEL_LOCAL: return shape_from_local(p, et);
        case JITR_B200_MODEL_WLH: return shape_from_wlh(p, et);
        case JITR_B200_MODEL_DOM: return shape_from_dom(p, et);
        default: return shape_from_canonical(p, et);
    }
}

}  // namespace jitr
