// C-ABI entry points of the elastic path: the sweep dispatcher (kernel template in
// elastic_sweep.cuh, one translation unit per nbasis) and the observable kernel
// (xs/elastic.py:310-381).  See include/jitr_b200.h for the contract of each entry point.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "common.cuh"
#include "jitr_b200.h"
#include "lu_warp.cuh"
#include "potentials.cuh"
#include "sweep_decl.cuh"

namespace jitr {

// ---------------------------------------------------------------------------------------------
// observables: one CTA per (energy, sample chunk); Legendre tables of that energy staged in
// shared memory once, then every sample of the chunk streams S in and dsdo/Ay/Q out, coalesced.
// ---------------------------------------------------------------------------------------------
struct ObsArgs {
    int lmax, n_angles, n_energies, chunk;
    int tables_in_smem;  // 0: the Legendre tables exceed shared memory (lmax >= 80 at 180 angles) and are read from global / L2
    long long n_samples;
    const double2* S;
    const int* nl;
    const double* P;
    const double* P1;
    const double2* phase;
    const double2* f_c;
    const double* kvec;
    double* dsdo;
    double* Ay;
    double* Q;
    double* tot_rxn;
    // optional weighted misfit against data (jitr_b200_elastic_misfit): sum_theta w (scale dsdo - y)^2 per (sample, energy)
    const double* data_y;
    const double* data_w;
    const double* data_scale;
    double* misfit;
};

constexpr int kObsWarps = 8;
constexpr int kObsAnglesPerLane = 8;  // angles a lane sums at once (register blocking over the l loop)

// One WARP per (sample, energy) pair, no CTA barrier in the sample loop: lanes l form the 2 L coefficient
// pairs of the sample (staged in the warp's slice of shared memory), then every lane sums up to eight
// angles at once over l, so a coefficient is fetched once per eight angles.
template <bool MISFIT>
__global__ void __launch_bounds__(kObsWarps * 32) elastic_observables_kernel(const ObsArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L1 = A.lmax + 1, NA = A.n_angles;
    const int e = blockIdx.y;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const double* Pg = A.P + (size_t)e * L1 * NA;
    const double* P1g = A.P1 + (size_t)e * L1 * NA;
    const double* Ps = Pg;    // [L1][NA]
    const double* P1s = P1g;  // [L1][NA]
    double2* coef = reinterpret_cast<double2*>(smem_raw);  // [warps][2][L1]
    if (A.tables_in_smem) {
        double* Pw = reinterpret_cast<double*>(smem_raw);
        double* P1w = Pw + (size_t)L1 * NA;
        coef = reinterpret_cast<double2*>(P1w + (size_t)L1 * NA);
        for (int i = threadIdx.x; i < L1 * NA; i += blockDim.x) {
            Pw[i] = Pg[i];
            P1w[i] = P1g[i];
        }
        Ps = Pw;
        P1s = P1w;
    }
    __syncthreads();
    double2* ca = coef + (size_t)warp * 2 * L1;  // coefficient of P_l
    double2* cb = ca + L1;                        // coefficient of P^1_l
    const double k = A.kvec[e];
    const long long s_begin = (long long)blockIdx.x * A.chunk;
    long long s_end = s_begin + A.chunk;
    if (s_end > A.n_samples) s_end = A.n_samples;
    for (long long s = s_begin + warp; s < s_end; s += kObsWarps) {
        const long long pair = s * A.n_energies + e;
        const int nl = A.nl[pair];
        const double2* Sp = A.S + (size_t)pair * 2 * L1;
        const double2* Sm = Sp + L1;
        double rx = 0.0, tt = 0.0;
        for (int l = lane; l < nl; l += 32) {
            const double2 sp = Sp[l], sm = Sm[l], ph = A.phase[(size_t)e * L1 + l];
            // (l+1)(S+ - 1) + l (S- - 1) ;  S+ - S-     (xs/elastic.py:357-362)
            const double2 wa = make_double2((l + 1) * (sp.x - 1.0) + l * (sm.x - 1.0), (l + 1) * sp.y + l * sm.y);
            const double2 wb = make_double2(sp.x - sm.x, sp.y - sm.y);
            ca[l] = cmul(ph, wa);
            cb[l] = cmul(ph, wb);
            rx += (l + 1) * (1.0 - (sp.x * sp.x + sp.y * sp.y)) + l * (1.0 - (sm.x * sm.x + sm.y * sm.y));
            tt += (l + 1) * (1.0 - sp.x) + l * (1.0 - sm.x);
        }
        for (int o = 16; o > 0; o >>= 1) {
            rx += __shfl_down_sync(0xffffffffu, rx, o);
            tt += __shfl_down_sync(0xffffffffu, tt, o);
        }
        if (lane == 0) {
            A.tot_rxn[2 * pair] = tt * (10.0 * 2.0 * M_PI / (k * k));
            A.tot_rxn[2 * pair + 1] = rx * (10.0 * M_PI / (k * k));
        }
        __syncwarp();
        double mis = 0.0;
        for (int t0 = 0; t0 < NA; t0 += 32 * kObsAnglesPerLane) {
            double2 fa[kObsAnglesPerLane], fb[kObsAnglesPerLane];
            int tj[kObsAnglesPerLane];
#pragma unroll
            for (int j = 0; j < kObsAnglesPerLane; ++j) {
                const int t = t0 + 32 * j + lane;
                tj[j] = t < NA ? t : NA - 1;
                fa[j] = A.f_c[(size_t)e * NA + tj[j]];
                fb[j] = make_double2(0.0, 0.0);
            }
            for (int l = 0; l < nl; ++l) {
                const double2 al = ca[l], bl = cb[l];
                const double* Pl = Ps + l * NA;
                const double* P1l = P1s + l * NA;
#pragma unroll
                for (int j = 0; j < kObsAnglesPerLane; ++j) {
                    if (t0 + 32 * j < NA) {  // warp-uniform
                        const double pl = Pl[tj[j]], p1 = P1l[tj[j]];
                        fa[j].x = fma(pl, al.x, fa[j].x);
                        fa[j].y = fma(pl, al.y, fa[j].y);
                        fb[j].x = fma(p1, bl.x, fb[j].x);
                        fb[j].y = fma(p1, bl.y, fb[j].y);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kObsAnglesPerLane; ++j) {
                const int t = t0 + 32 * j + lane;
                if (t < NA) {
                    const double d0 = fa[j].x * fa[j].x + fa[j].y * fa[j].y + fb[j].x * fb[j].x + fb[j].y * fb[j].y;
                    const double den = fmax(d0, 1e-30);
                    // conj(a) b = (a.x - i a.y)(b.x + i b.y)
                    const double re = fa[j].x * fb[j].x + fa[j].y * fb[j].y, im = fa[j].x * fb[j].y - fa[j].y * fb[j].x;
                    const size_t o = (size_t)pair * NA + t;
                    if (!MISFIT || A.dsdo) A.dsdo[o] = d0 * 10.0;
                    if (A.Ay) A.Ay[o] = 2.0 * im / den;
                    if (A.Q) A.Q[o] = 2.0 * re / den;
                    if (MISFIT) {
                        const size_t d = (size_t)e * NA + t;
                        const double v = d0 * 10.0 * (A.data_scale ? A.data_scale[d] : 1.0) - A.data_y[d];
                        mis = fma(A.data_w[d] * v, v, mis);
                    }
                }
            }
        }
        if (MISFIT) {
            for (int o = 16; o > 0; o >>= 1) mis += __shfl_down_sync(0xffffffffu, mis, o);
            if (lane == 0) A.misfit[pair] = mis;
        }
        __syncwarp();  // the coefficients of this sample are consumed before the next one overwrites them
    }
}


// ---------------------------------------------------------------------------------------------
// The same observables as a GEMM on the FP64 tensor path (round 2).  a(theta) = f_c + sum_l P_l(theta) ca_l and
// b(theta) = sum_l P1_l(theta) cb_l are two real GEMMs per group of FOUR samples: rows = (sample, re | im) of the complex
// coefficients, K = l, N = theta -- DMMA m8n8k4 with the Legendre tables as the B operand from shared memory (row stride = 4 mod 16:
// the fragment loads are bank-conflict free) and the coefficient rows as the A operand (staged per warp, kept in registers across the
// angle tiles).  A table entry is fetched once per four samples instead of once per sample, and the k loop stops at the largest
// number of partial waves kept in the group.  Epilogue: the (re, im) rows of a sample sit four lanes apart -> one exchange, then
// every lane writes one (sample, angle).  Sums run in the DMMA's order: same terms, rounding-level differences only.
// ---------------------------------------------------------------------------------------------
constexpr int kObsGroup = 4;  // samples per DMMA row block (8 rows = 4 samples x (re, im))
constexpr int kObsDmmaWarps = 16;  // one CTA per SM (the tables take 92 KB): sixteen warps hide the fragment-load and DMMA latencies

template <bool MISFIT>
__global__ void __launch_bounds__(kObsDmmaWarps * 32) elastic_observables_dmma_kernel(const ObsArgs A, const int Lp, const int ts,
                                                                                         const int chunks_per_energy) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L1 = A.lmax + 1, NA = A.n_angles;
    double* Ps = reinterpret_cast<double*>(smem_raw);   // [Lp][ts], rows >= L1 and columns >= NA zero
    double* P1s = Ps + (size_t)Lp * ts;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    // per warp: coefficient rows [2 GEMMs][8 rows][Lp + 1]
    const int cs = Lp + 1;
    double* coef = P1s + (size_t)Lp * ts + (size_t)warp * 2 * 8 * cs;
    const int fr = lane >> 2, fk = lane & 3;       // accumulator row (sample fr / 2, part fr & 1), column pair 2 fk, 2 fk + 1
    const int my_s = fr >> 1, my_part = fr & 1;
    // persistent CTAs over the (energy, sample chunk) items, energy-major and contiguous per CTA: a CTA reloads its tables only
    // when its next item belongs to another energy (at most once or twice), and the grid is exactly one CTA per SM
    const long long n_items = (long long)A.n_energies * chunks_per_energy;
    const long long per_cta = (n_items + gridDim.x - 1) / gridDim.x;
    long long item = (long long)blockIdx.x * per_cta;
    const long long item_end = item + per_cta < n_items ? item + per_cta : n_items;
    int e_loaded = -1;
    for (; item < item_end; ++item) {
    const int e = (int)(item / chunks_per_energy);
    if (e != e_loaded) {
        __syncthreads();
        const double* Pg = A.P + (size_t)e * L1 * NA;
        const double* P1g = A.P1 + (size_t)e * L1 * NA;
        for (int i = threadIdx.x; i < Lp * ts; i += blockDim.x) {
            const int l = i / ts, t = i - l * ts;
            const bool in = l < L1 && t < NA;
            Ps[i] = in ? Pg[l * NA + t] : 0.0;
            P1s[i] = in ? P1g[l * NA + t] : 0.0;
        }
        e_loaded = e;
        __syncthreads();
    }
    const double k = A.kvec[e];
    const long long s_begin = (item - (long long)e * chunks_per_energy) * A.chunk;
    long long s_end = s_begin + A.chunk;
    if (s_end > A.n_samples) s_end = A.n_samples;
    for (long long s0 = s_begin + (long long)warp * kObsGroup; s0 < s_end; s0 += (long long)kObsDmmaWarps * kObsGroup) {
        // ---- coefficients of the four samples: lane = l (two rounds cover Lp <= 64)
        int nlmax = 0;
#pragma unroll
        for (int g = 0; g < kObsGroup; ++g) {
            const long long smp = s0 + g;
            const bool live = smp < s_end;
            const long long pair = (live ? smp : s_begin) * A.n_energies + e;
            const int nl = live ? A.nl[pair] : 0;
            nlmax = max(nlmax, nl);
            const double2* Sp = A.S + (size_t)pair * 2 * L1;
            const double2* Sm = Sp + L1;
            double rx = 0.0, tt = 0.0;
            for (int l = lane; l < Lp; l += 32) {
                double2 wa = make_double2(0.0, 0.0), wb = wa;
                if (l < nl) {
                    const double2 sp = Sp[l], sm = Sm[l], ph = A.phase[(size_t)e * L1 + l];
                    // (l+1)(S+ - 1) + l (S- - 1) ;  S+ - S-     (xs/elastic.py:357-362)
                    wa = cmul(ph, make_double2((l + 1) * (sp.x - 1.0) + l * (sm.x - 1.0), (l + 1) * sp.y + l * sm.y));
                    wb = cmul(ph, make_double2(sp.x - sm.x, sp.y - sm.y));
                    rx += (l + 1) * (1.0 - (sp.x * sp.x + sp.y * sp.y)) + l * (1.0 - (sm.x * sm.x + sm.y * sm.y));
                    tt += (l + 1) * (1.0 - sp.x) + l * (1.0 - sm.x);
                }
                coef[(2 * g) * cs + l] = wa.x;
                coef[(2 * g + 1) * cs + l] = wa.y;
                coef[(8 + 2 * g) * cs + l] = wb.x;
                coef[(8 + 2 * g + 1) * cs + l] = wb.y;
            }
            for (int o = 16; o > 0; o >>= 1) {
                rx += __shfl_down_sync(0xffffffffu, rx, o);
                tt += __shfl_down_sync(0xffffffffu, tt, o);
            }
            if (lane == 0 && live) {
                A.tot_rxn[2 * pair] = tt * (10.0 * 2.0 * M_PI / (k * k));
                A.tot_rxn[2 * pair + 1] = rx * (10.0 * M_PI / (k * k));
            }
        }
        __syncwarp();
        const int ksteps = (nlmax + 3) >> 2;  // warp-uniform
        // A fragments (row fr, k = fk) of both GEMMs for every k step, in registers across the angle tiles
        double aa[16], ab[16];
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
            aa[ks] = ks < ksteps ? coef[fr * cs + 4 * ks + fk] : 0.0;
            ab[ks] = ks < ksteps ? coef[(8 + fr) * cs + 4 * ks + fk] : 0.0;
        }
        const long long my_smp = s0 + my_s;
        const bool my_live = my_smp < s_end;
        const long long my_pair = (my_live ? my_smp : s_begin) * A.n_energies + e;
        double mis = 0.0;
        // four angle tiles at a time: eight independent accumulator chains per lane hide the DMMA latency
        constexpr int TU = 4;
        for (int tb = 0; tb < NA; tb += 8 * TU) {
            double ca0[TU], ca1[TU], cb0[TU], cb1[TU];
#pragma unroll
            for (int u = 0; u < TU; ++u) {
                const int tc = tb + 8 * u + 2 * fk;  // this lane's accumulator columns tc, tc + 1 of tile u
                // a starts from the Coulomb amplitude (its re / im part on the re / im rows), b from zero
                const double2 f0 = A.f_c[(size_t)e * NA + (tc < NA ? tc : NA - 1)], f1 = A.f_c[(size_t)e * NA + (tc + 1 < NA ? tc + 1 : NA - 1)];
                ca0[u] = my_part ? f0.y : f0.x;
                ca1[u] = my_part ? f1.y : f1.x;
                cb0[u] = cb1[u] = 0.0;
            }
            const double* Pb = Ps + fk * ts + tb + fr;    // B fragment: row k = fk, column n = fr
            const double* P1b = P1s + fk * ts + tb + fr;
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                if (ks < ksteps) {
#pragma unroll
                    for (int u = 0; u < TU; ++u) {
                        if (tb + 8 * u < NA) {  // warp-uniform
                            dmma884(ca0[u], ca1[u], aa[ks], Pb[4 * ks * ts + 8 * u]);
                            dmma884(cb0[u], cb1[u], ab[ks], P1b[4 * ks * ts + 8 * u]);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < TU; ++u) {
                if (tb + 8 * u >= NA) break;
                const int tc = tb + 8 * u + 2 * fk;
                // the other part (re <-> im) of the same sample and angles lives four lanes away
                const double oa0 = __shfl_xor_sync(0xffffffffu, ca0[u], 4), oa1 = __shfl_xor_sync(0xffffffffu, ca1[u], 4);
                const double ob0 = __shfl_xor_sync(0xffffffffu, cb0[u], 4), ob1 = __shfl_xor_sync(0xffffffffu, cb1[u], 4);
                // the re lane writes angle tc, the im lane angle tc + 1
                const double are = my_part ? oa1 : ca0[u], aim = my_part ? ca1[u] : oa0;
                const double bre = my_part ? ob1 : cb0[u], bim = my_part ? cb1[u] : ob0;
                const int t = tc + my_part;
                if (my_live && t < NA) {
                    const double d0 = are * are + aim * aim + bre * bre + bim * bim;
                    const double den = fmax(d0, 1e-30);
                    const double re = are * bre + aim * bim, im = are * bim - aim * bre;  // conj(a) b
                    const size_t o = (size_t)my_pair * NA + t;
                    if (!MISFIT || A.dsdo) A.dsdo[o] = d0 * 10.0;
                    if (A.Ay) A.Ay[o] = 2.0 * im / den;
                    if (A.Q) A.Q[o] = 2.0 * re / den;
                    if (MISFIT) {
                        const size_t d = (size_t)e * NA + t;
                        const double v = d0 * 10.0 * (A.data_scale ? A.data_scale[d] : 1.0) - A.data_y[d];
                        mis = fma(A.data_w[d] * v, v, mis);
                    }
                }
            }
        }
        if (MISFIT) {
            // the eight lanes 8 s .. 8 s + 7 hold the terms of sample s
            for (int o = 4; o > 0; o >>= 1) mis += __shfl_xor_sync(0xffffffffu, mis, o);
            if ((lane & 7) == 0 && my_live) A.misfit[my_pair] = mis;
        }
        __syncwarp();  // the coefficient rows are consumed before the next group overwrites them
    }
    }  // items
}

// ---------------------------------------------------------------------------------------------
// stand-alone evaluation of the built-in potential forms on the mesh (same device code as the
// fused sweep): SingleChannelOpticalModel.evaluate, omp.py:26-43 / kduq.py:541-559
// ---------------------------------------------------------------------------------------------
__global__ void eval_potentials_kernel(int N, int model, long long n_pairs, int n_energies, const double* mesh_x,
                                       const double* etab, const double* params, int n_params, double2* central,
                                       double2* spin_orbit, double* coulomb) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pairs * N) return;
    const long long pair = idx / N;
    const int n = (int)(idx - pair * N);
    const long long smp = pair / n_energies;
    const int e = (int)(pair - smp * n_energies);
    const double* et = etab + (size_t)e * JITR_B200_ET_STRIDE;
    const size_t row = model == JITR_B200_MODEL_SHAPE_PAIR ? (size_t)pair : (size_t)smp;
    const Shape sh = shape_from_model(model, params + row * n_params, et);
    double2 vc, vs;
    double co;
    eval_shape(sh, mesh_x[n] * et[JITR_B200_ET_RCH], vc, vs, co);
    central[idx] = vc;
    spin_orbit[idx] = vs;
    coulomb[idx] = co;
}

// pick the solver mode for nbasis: fast path when P <= 40, generic otherwise
template <bool PARAMS>
static int launch_sweep(const SweepArgs& a, cudaStream_t st) {
    const int N = a.N;
    if (N < 1 || N > 64) return set_error(JITR_B200_ERR_UNSUPPORTED, "elastic sweep: nbasis must be in 1..64");
    const int P = (N + 7) & ~7;
    switch (P) {
        case 16: return launch_sweep_mode<PARAMS, 16>(a, st);
        case 24: return launch_sweep_mode<PARAMS, 24>(a, st);
        case 32: return launch_sweep_mode<PARAMS, 32>(a, st);
#ifndef JITR_PACKED40
        case 40: return launch_sweep_mode<PARAMS, 40>(a, st);
#endif
        default: break;
    }
#ifdef JITR_PACKED40
    constexpr int kPackedFrom = 40;
#else
    constexpr int kPackedFrom = 48;
#endif
    if (P >= kPackedFrom && P <= 64 && a.symmetric && a.n_pairs < 0x7fffffffLL && !a.pair_list) {
        // packed modes: lower triangle only, six systems per SM at P = 64; the pairs they defer (a system failed the
        // threshold test) are redone by the partial-pivoting modes right behind, on the same stream
        int* buf = defer_buffer(a.n_pairs, st);
        if (!buf) return set_cuda_error("elastic sweep: deferred-pairs buffer");
        SweepArgs p = a;
        p.defer_count = buf;
        p.defer_list = buf + 4;
        int rc = JITR_B200_ERR_UNSUPPORTED;
        switch (P) {
#ifdef JITR_PACKED40
            case 40: rc = launch_sweep_mode<PARAMS, 40>(p, st); break;
#endif
            case 48: rc = launch_sweep_mode<PARAMS, 48>(p, st); break;
            case 56: rc = launch_sweep_mode<PARAMS, 56>(p, st); break;
            case 64: rc = launch_sweep_mode<PARAMS, 64>(p, st); break;
            default: break;
        }
        if (rc != JITR_B200_OK) return rc;
        SweepArgs r = a;
        r.pair_list = buf + 4;
        r.pair_count = buf;
        return P + 1 <= 64 ? launch_sweep_mode<PARAMS, -2>(r, st) : launch_sweep_mode<PARAMS, -3>(r, st);
    }
    if (P + 1 <= 64) return launch_sweep_mode<PARAMS, -2>(a, st);
    return launch_sweep_mode<PARAMS, -3>(a, st);
}

}  // namespace jitr

using namespace jitr;

extern "C" int jitr_b200_elastic_sweep_params(int nbasis, int lmax, int model, long long n_samples, int n_energies,
                                              const double* That, const double* mesh_x, const double* bvec,
                                              const double* etab, const double* asym, const double* params,
                                              int n_params, double tol, double* S_out, int* nl_out, int* status,
                                              void* stream) {
    if (n_samples == 0) return JITR_B200_OK;  // empty batch (its tensors have null data pointers)
    if (lmax < 0 || n_samples < 0 || n_energies < 1 || !That || !mesh_x || !bvec || !etab || !asym || !params ||
        !S_out || !nl_out || !status)
        return set_error(JITR_B200_ERR_BAD_ARG, "elastic_sweep_params: null pointer or bad size");
    const int need = model == JITR_B200_MODEL_KDUQ ? 40 : model == JITR_B200_MODEL_CHUQ ? 22 : model == JITR_B200_MODEL_LOCAL ? 15 : model == JITR_B200_MODEL_WLH ? 46 : model == JITR_B200_MODEL_SHAPE_PAIR ? 18 : model == JITR_B200_MODEL_DOM ? 19 : 17;
    if (model < 0 || model > 6 || n_params < need)
        return set_error(JITR_B200_ERR_BAD_ARG, "elastic_sweep_params: model / n_params mismatch");
    SweepArgs a{};
    a.lmax = lmax; a.model = model; a.n_energies = n_energies; a.n_params = n_params;
    a.n_pairs = n_samples * n_energies;
    a.That = That; a.mesh_x = mesh_x; a.bvec = bvec; a.etab = etab;
    a.asym = reinterpret_cast<const double2*>(asym);
    a.params = params; a.tol = tol;
    a.S_out = reinterpret_cast<double2*>(S_out); a.nl_out = nl_out; a.status = status;
    a.N = nbasis;
    a.symmetric = option(JITR_B200_OPT_SYMMETRIC);
    a.fallbacks = a.symmetric ? fallback_counter() : nullptr;
    a.fallback_mask = a.symmetric ? fallback_mask() : nullptr;
    return launch_sweep<true>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int jitr_b200_elastic_sweep_tensors(int nbasis, int lmax, long long n_samples, int n_energies,
                                               const double* That, const double* mesh_x, const double* bvec,
                                               const double* etab, const double* asym, const double* central,
                                               const double* spin_orbit, const double* coulomb, double tol,
                                               double* S_out, int* nl_out, int* status, void* stream) {
    if (n_samples == 0) return JITR_B200_OK;  // empty batch (its tensors have null data pointers)
    if (lmax < 0 || n_samples < 0 || n_energies < 1 || !That || !mesh_x || !bvec || !etab || !asym || !central ||
        !S_out || !nl_out || !status)
        return set_error(JITR_B200_ERR_BAD_ARG, "elastic_sweep_tensors: null pointer or bad size");
    SweepArgs a{};
    a.lmax = lmax; a.model = 0; a.n_energies = n_energies; a.n_params = 0;
    a.n_pairs = n_samples * n_energies;
    a.That = That; a.mesh_x = mesh_x; a.bvec = bvec; a.etab = etab;
    a.asym = reinterpret_cast<const double2*>(asym);
    a.central = reinterpret_cast<const double2*>(central);
    a.spin_orbit = reinterpret_cast<const double2*>(spin_orbit);
    a.coulomb = reinterpret_cast<const double2*>(coulomb);
    a.tol = tol;
    a.S_out = reinterpret_cast<double2*>(S_out); a.nl_out = nl_out; a.status = status;
    a.N = nbasis;
    a.symmetric = option(JITR_B200_OPT_SYMMETRIC);
    a.fallbacks = a.symmetric ? fallback_counter() : nullptr;
    a.fallback_mask = a.symmetric ? fallback_mask() : nullptr;
    return launch_sweep<false>(a, static_cast<cudaStream_t>(stream));
}

static int launch_observables(ObsArgs a, const char* what, cudaStream_t st) {
    const int L1 = a.lmax + 1;
    // GEMM formulation on DMMA when the zero-padded tables fit shared memory (lmax <= 63)
    {
        const int Lp = (L1 + 3) & ~3;
        int ts = (a.n_angles + 7) & ~7;
        while ((ts & 15) != 4) ++ts;  // row stride = 4 mod 16: conflict-free fragment loads
        const size_t smem = (size_t)2 * Lp * ts * 8 + (size_t)kObsDmmaWarps * 2 * 8 * (Lp + 1) * 8;
        if (Lp <= 64 && smem <= 232448 && option(JITR_B200_OPT_OBSERVABLES_DMMA)) {
            auto kern = a.misfit ? elastic_observables_dmma_kernel<true> : elastic_observables_dmma_kernel<false>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return set_cuda_error("cudaFuncSetAttribute(observables)");
            // items = (energy, chunk of samples); about sixteen items per CTA so that the contiguous split over the SMs is even
            const int sms = sm_count();
            const long long per = kObsDmmaWarps * kObsGroup;  // samples one pass of the CTA covers
            long long chunks = (16LL * sms + a.n_energies - 1) / a.n_energies;
            const long long max_chunks = (a.n_samples + per - 1) / per;
            if (chunks > max_chunks) chunks = max_chunks;
            if (chunks < 1) chunks = 1;
            long long chunk = (a.n_samples + chunks - 1) / chunks;
            chunk = (chunk + per - 1) / per * per;
            chunks = (a.n_samples + chunk - 1) / chunk;
            a.chunk = (int)(chunk > (1 << 30) ? (1 << 30) : chunk);
            long long blocks = (long long)a.n_energies * chunks;
            if (blocks > sms) blocks = sms;
            kern<<<(unsigned)blocks, kObsDmmaWarps * 32, smem, st>>>(a, Lp, ts, (int)chunks);
            count_launch();
            if (cudaGetLastError() != cudaSuccess) return set_cuda_error(what);
            return JITR_B200_OK;
        }
    }
    size_t smem = (size_t)2 * L1 * a.n_angles * 8 + (size_t)kObsWarps * 2 * L1 * 16;
    a.tables_in_smem = 1;
    if (smem > 232448) {  // large lmax (the Frescox regression cases reach 100): tables from global memory / L2
        a.tables_in_smem = 0;
        smem = (size_t)kObsWarps * 2 * L1 * 16;
    }
    auto kern = a.misfit ? elastic_observables_kernel<true> : elastic_observables_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(observables)");
    // chunk the sample axis so that the grid is a few waves of the SM count
    const int sms = sm_count();
    long long target_blocks_x = (4LL * sms + a.n_energies - 1) / a.n_energies;
    if (target_blocks_x < 1) target_blocks_x = 1;
    long long chunk = (a.n_samples + target_blocks_x - 1) / target_blocks_x;
    if (chunk < 1) chunk = 1;
    a.chunk = (int)(chunk > (1 << 30) ? (1 << 30) : chunk);
    dim3 grid((unsigned)((a.n_samples + a.chunk - 1) / a.chunk), (unsigned)a.n_energies);
    kern<<<grid, kObsWarps * 32, smem, st>>>(a);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error(what);
    return JITR_B200_OK;
}

extern "C" int jitr_b200_elastic_observables(int lmax, int n_angles, long long n_samples, int n_energies,
                                             const double* S, const int* nl, const double* P, const double* P1,
                                             const double* phase, const double* f_c, const double* kvec,
                                             double* dsdo, double* Ay, double* Q, double* tot_rxn, void* stream) {
    if (n_samples == 0) return JITR_B200_OK;  // empty batch (its tensors have null data pointers)
    if (lmax < 0 || lmax > 255 || n_angles < 1 || n_samples < 0 || n_energies < 1 || !S || !nl || !P || !P1 || !phase ||
        !f_c || !kvec || !dsdo || !tot_rxn)
        return set_error(JITR_B200_ERR_BAD_ARG, "elastic_observables: null pointer or bad size");
    ObsArgs a{};
    a.lmax = lmax; a.n_angles = n_angles; a.n_energies = n_energies; a.n_samples = n_samples;
    a.S = reinterpret_cast<const double2*>(S); a.nl = nl; a.P = P; a.P1 = P1;
    a.phase = reinterpret_cast<const double2*>(phase);
    a.f_c = reinterpret_cast<const double2*>(f_c);
    a.kvec = kvec; a.dsdo = dsdo; a.Ay = Ay; a.Q = Q; a.tot_rxn = tot_rxn;
    return launch_observables(a, "elastic_observables launch", static_cast<cudaStream_t>(stream));
}

extern "C" int jitr_b200_elastic_misfit(int lmax, int n_angles, long long n_samples, int n_energies, const double* S,
                                        const int* nl, const double* P, const double* P1, const double* phase,
                                        const double* f_c, const double* kvec, const double* y, const double* w,
                                        const double* scale, double* misfit, double* tot_rxn, double* dsdo, void* stream) {
    if (n_samples == 0) return JITR_B200_OK;
    if (lmax < 0 || lmax > 255 || n_angles < 1 || n_samples < 0 || n_energies < 1 || !S || !nl || !P || !P1 || !phase ||
        !f_c || !kvec || !y || !w || !misfit || !tot_rxn)
        return set_error(JITR_B200_ERR_BAD_ARG, "elastic_misfit: null pointer or bad size");
    ObsArgs a{};
    a.lmax = lmax; a.n_angles = n_angles; a.n_energies = n_energies; a.n_samples = n_samples;
    a.S = reinterpret_cast<const double2*>(S); a.nl = nl; a.P = P; a.P1 = P1;
    a.phase = reinterpret_cast<const double2*>(phase);
    a.f_c = reinterpret_cast<const double2*>(f_c);
    a.kvec = kvec; a.dsdo = dsdo; a.tot_rxn = tot_rxn;
    a.data_y = y; a.data_w = w; a.data_scale = scale; a.misfit = misfit;
    return launch_observables(a, "elastic_misfit launch", static_cast<cudaStream_t>(stream));
}

extern "C" int jitr_b200_eval_potentials(int nbasis, int model, long long n_samples, int n_energies, const double* mesh_x,
                                         const double* etab, const double* params, int n_params, double* central,
                                         double* spin_orbit, double* coulomb, void* stream) {
    if (n_samples == 0) return JITR_B200_OK;  // empty batch (its tensors have null data pointers)
    if (nbasis < 1 || n_samples < 0 || n_energies < 1 || !mesh_x || !etab || !params || !central || !spin_orbit || !coulomb)
        return set_error(JITR_B200_ERR_BAD_ARG, "eval_potentials: null pointer or bad size");
    const int need = model == JITR_B200_MODEL_KDUQ ? 40 : model == JITR_B200_MODEL_CHUQ ? 22 : model == JITR_B200_MODEL_LOCAL ? 15 : model == JITR_B200_MODEL_WLH ? 46 : model == JITR_B200_MODEL_SHAPE_PAIR ? 18 : model == JITR_B200_MODEL_DOM ? 19 : 17;
    if (model < 0 || model > 6 || n_params < need) return set_error(JITR_B200_ERR_BAD_ARG, "eval_potentials: model / n_params mismatch");
    const long long total = n_samples * n_energies * nbasis;
    if (total == 0) return JITR_B200_OK;
    eval_potentials_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        nbasis, model, n_samples * n_energies, n_energies, mesh_x, etab, params, n_params,
        reinterpret_cast<double2*>(central), reinterpret_cast<double2*>(spin_orbit), coulomb);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("eval_potentials launch");
    return JITR_B200_OK;
}
