// Fused spin-1/2 elastic sweep kernel:
// potential evaluation -> system assembly -> bordered LU -> R -> S, one WARP per (sample, energy)
// pair walking l = 0..lmax, j = l +- 1/2 with the reference's data-dependent early break
// (xs/elastic.py:104-183).  nbasis is a runtime value (<= 63: at most two row slots per lane).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"
#include "jitr_b200.h"
#include "lu_warp.cuh"
#include "potentials.cuh"
#include "sweep_decl.cuh"

namespace jitr {

constexpr int kSweepMaxSmem = 232448;  // max dynamic shared memory per block on sm_100
constexpr int kSweepMaxWarps = 8;

// shared-memory plan for a given nbasis: `warps` bordered systems of (P+1)^2 complex + That (N^2 real)
struct SweepPlan {
    int P, LD, msz, warps, smem;
};
inline SweepPlan sweep_plan(int N) {
    SweepPlan p;
    p.P = (N + 7) & ~7;
    p.LD = p.P + 1;
    p.msz = p.LD * p.LD;
    const int fit = (kSweepMaxSmem - N * N * 8) / (p.msz * 16);
    p.warps = fit > kSweepMaxWarps ? kSweepMaxWarps : fit;
    p.smem = p.warps * p.msz * 16 + N * N * 8;
    return p;
}

// PARAMS: built-in potential forms from parameter rows; otherwise mesh values from tensors.
template <bool PARAMS>
__global__ void __launch_bounds__(kSweepMaxWarps * 32, 1) elastic_sweep_kernel(const SweepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = A.N;
    LuGeom g;
    g.N = N;
    g.P = (N + 7) & ~7;
    g.LD = g.P + 1;
    const int P = g.P, LD = g.LD, msz = LD * LD;
    const int nwarps = blockDim.x >> 5;
    double2* Mall = reinterpret_cast<double2*>(smem_raw);
    double* Ts = reinterpret_cast<double*>(smem_raw + (size_t)nwarps * msz * 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* M = Mall + (size_t)warp * msz;

    for (int i = threadIdx.x; i < N * N; i += blockDim.x) Ts[i] = A.That[i];
    __syncthreads();

    // mesh points owned by this lane (two slots, N <= 63)
    const int n0 = lane, n1 = lane + 32;
    const bool v1 = n1 < N;
    const bool v0 = n0 < N;
    const double x0 = A.mesh_x[v0 ? n0 : 0], x1 = A.mesh_x[v1 ? n1 : 0];
    const double b0 = A.bvec[v0 ? n0 : 0], b1 = A.bvec[v1 ? n1 : 0];
    const double t0 = Ts[(v0 ? n0 : 0) * (N + 1)], t1 = Ts[(v1 ? n1 : 0) * (N + 1)];
    const double cent0 = 1.0 / (x0 * x0), cent1 = 1.0 / (x1 * x1);
    const int L1 = A.lmax + 1;
    const double2 zero2 = make_double2(0.0, 0.0);

    for (long long pair = (long long)blockIdx.x * nwarps + warp; pair < A.n_pairs;
         pair += (long long)gridDim.x * nwarps) {
        const long long smp = pair / A.n_energies;
        const int e = (int)(pair - smp * A.n_energies);
        const double* et = A.etab + (size_t)e * JITR_B200_ET_STRIDE;
        const double a = et[JITR_B200_ET_A], E0 = et[JITR_B200_ET_ECM], radius = et[JITR_B200_ET_RCH];
        const double a2 = a * a;

        // ---- potentials on this lane's mesh points, in the a^2-scaled form of the diagonal
        double2 vc0, vs0, vc1, vs1;
        if (PARAMS) {
            const Shape sh = shape_from_model(A.model, A.params + (size_t)smp * A.n_params, et);
            double co0, co1;
            eval_shape(sh, x0 * radius, vc0, vs0, co0);
            eval_shape(sh, x1 * radius, vc1, vs1, co1);
            vc0.x += co0;
            vc1.x += co1;
        } else {
            const size_t off = (size_t)pair * N;
            vc0 = A.central[off + (v0 ? n0 : 0)];
            vc1 = A.central[off + (v1 ? n1 : 0)];
            vs0 = vs1 = zero2;
            if (A.spin_orbit) {
                vs0 = A.spin_orbit[off + (v0 ? n0 : 0)];
                vs1 = A.spin_orbit[off + (v1 ? n1 : 0)];
            }
            if (A.coulomb) {
                const double2 c0 = A.coulomb[off + (v0 ? n0 : 0)], c1 = A.coulomb[off + (v1 ? n1 : 0)];
                vc0.x += c0.x; vc0.y += c0.y;
                vc1.x += c1.x; vc1.y += c1.y;
            }
        }
        // diag(a^2 A) = That_nn + l(l+1)/x_n^2 + a^2(-1 + V/E0) + (l.s) a^2 Vso/E0
        const double2 d0 = make_double2(t0 + a2 * (vc0.x / E0 - 1.0), a2 * (vc0.y / E0));
        const double2 d1 = make_double2(t1 + a2 * (vc1.x / E0 - 1.0), a2 * (vc1.y / E0));
        const double2 so0 = make_double2(a2 * (vs0.x / E0), a2 * (vs0.y / E0));
        const double2 so1 = make_double2(a2 * (vs1.x / E0), a2 * (vs1.y / E0));

        double2* Sp_out = A.S_out + (size_t)pair * 2 * L1;
        double2* Sm_out = Sp_out + L1;
        int nl = 0, singular = 0, nonfinite = 0;
        for (int l = 0; l <= A.lmax; ++l) {
            const double ll1 = (double)l * (double)(l + 1);
            const double2* as = A.asym + ((size_t)e * L1 + l) * 4;
            const double2 Hp = as[0], Hm = as[1], Hpp = as[2], Hmp = as[3];
            double2 Sj[2];
            Sj[1] = zero2;
            const int nj = (l == 0) ? 1 : 2;
            for (int jj = 0; jj < nj; ++jj) {
                const double ls = (jj == 0) ? (double)l : -(double)(l + 1);
                // ---- assemble the bordered matrix  [[That + diag, 0, b], [0, 0, 0], [b^T, 0, 0]]
                for (int i = 0; i < N; ++i) {
                    const double* Trow = Ts + i * N;
                    double2* Mrow = M + i * LD;
                    if (lane < P) Mrow[lane] = make_double2(v0 ? Trow[n0] : 0.0, 0.0);
                    if (n1 < P) Mrow[n1] = make_double2(v1 ? Trow[n1] : 0.0, 0.0);
                }
                for (int i = N; i < P; ++i) {  // zero padding rows
                    for (int j = lane; j < LD; j += 32) M[i * LD + j] = zero2;
                }
                __syncwarp();
                if (v0) {
                    M[n0 * LD + n0] = make_double2(d0.x + ll1 * cent0 + ls * so0.x, d0.y + ls * so0.y);
                    M[n0 * LD + P] = make_double2(b0, 0.0);
                    M[P * LD + n0] = make_double2(b0, 0.0);
                }
                if (v1) {
                    M[n1 * LD + n1] = make_double2(d1.x + ll1 * cent1 + ls * so1.x, d1.y + ls * so1.y);
                    M[n1 * LD + P] = make_double2(b1, 0.0);
                    M[P * LD + n1] = make_double2(b1, 0.0);
                }
                for (int j = N + lane; j <= P; j += 32) M[P * LD + j] = zero2;  // z row: padding + corner
                __syncwarp();
                // ---- factorise; corner = -b^T (a^2 A)^-1 b = -R
                const double2 corner = bordered_schur(M, g, lane, singular);
                __syncwarp();
                const double2 aR = make_double2(-a * corner.x, -a * corner.y);
                // S = (H- - a R H-') / (H+ - a R H+')   (rmatrix/core.py:51-56 with nch = 1)
                const double2 t_m = cmul(aR, Hmp), t_p = cmul(aR, Hpp);
                const double2 num = make_double2(Hm.x - t_m.x, Hm.y - t_m.y);
                const double2 den = make_double2(Hp.x - t_p.x, Hp.y - t_p.y);
                const double2 S = cdiv(num, den);
                Sj[jj] = S;
                if (!(isfinite(S.x) && isfinite(S.y))) nonfinite = 1;
            }
            if (lane == 0) {
                Sp_out[l] = Sj[0];
                Sm_out[l] = Sj[1];
            }
            nl = l + 1;
            if (l >= 1 && A.tol > 0.0) {
                const double dp = hypot(1.0 - Sj[0].x, Sj[0].y), dm = hypot(1.0 - Sj[1].x, Sj[1].y);
                if (dp < A.tol && dm < A.tol) break;
            }
        }
        for (int l = nl + lane; l <= A.lmax; l += 32) {
            Sp_out[l] = zero2;
            Sm_out[l] = zero2;
        }
        if (lane == 0) {
            A.nl_out[pair] = nl;
            A.status[pair] = nonfinite ? JITR_B200_STATUS_NONFINITE : (singular ? JITR_B200_STATUS_SINGULAR : JITR_B200_STATUS_OK);
        }
    }
}

template <bool PARAMS>
int launch_sweep(const SweepArgs& a, cudaStream_t st) {
    if (a.N < 1 || a.N > 63) return set_error(JITR_B200_ERR_UNSUPPORTED, "elastic sweep: nbasis must be in 1..63");
    const SweepPlan plan = sweep_plan(a.N);
    if (plan.warps < 1) return set_error(JITR_B200_ERR_UNSUPPORTED, "elastic sweep: system does not fit in shared memory");
    auto kern = elastic_sweep_kernel<PARAMS>;
    static int configured = 0;
    if (plan.smem > configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem) != cudaSuccess)
            return set_cuda_error("cudaFuncSetAttribute(elastic_sweep)");
        configured = plan.smem;
    }
    const int sms = sm_count();
    long long blocks = (a.n_pairs + plan.warps - 1) / plan.warps;
    if (blocks > sms) blocks = sms;
    if (blocks < 1) return JITR_B200_OK;
    kern<<<(unsigned)blocks, plan.warps * 32, plan.smem, st>>>(a);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("elastic_sweep launch");
    return JITR_B200_OK;
}

}  // namespace jitr
