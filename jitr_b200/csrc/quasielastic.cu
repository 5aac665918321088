// Quasi-elastic (p,n) in the DWBA (xs/quasielastic_pn.py): the two steps that follow the distorted-wave
// solves (jitr_b200_solve_batched with coefficients) -- the radial overlap T_lj and the angular sum.
#include <cuda_runtime.h>

#include "common.cuh"
#include "jitr_b200.h"

namespace jitr {

constexpr int kQeWarps = 8;

// T_lj = sum_n x_p[n] (U1c[n] + (l.s)_j U1so[n]) x_n[n] * scale     (quasielastic_pn.py:317-322)
// one warp per (j, sample); lanes stride the mesh
__global__ void __launch_bounds__(kQeWarps * 32) qe_overlap_kernel(int N, long long batch, int nj, const double2* __restrict__ xp,
                                                                   const double2* __restrict__ xn,
                                                                   const double2* __restrict__ U1c,
                                                                   const double2* __restrict__ U1so, double lds0, double lds1,
                                                                   double scale, double2* __restrict__ out,
                                                                   long long out_sample_stride, long long out_j_stride) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * kQeWarps + (threadIdx.x >> 5);
    if (w >= batch * nj) return;
    const int j = (int)(w / batch);
    const long long b = w - (long long)j * batch;
    const double lds = j == 0 ? lds0 : lds1;
    const double2* p = xp + (size_t)w * N;
    const double2* q = xn + (size_t)w * N;
    const double2* uc = U1c + (size_t)b * N;
    const double2* us = U1so ? U1so + (size_t)b * N : nullptr;
    double ax = 0.0, ay = 0.0;
    for (int n = lane; n < N; n += 32) {
        double2 u = uc[n];
        if (us) {
            const double2 s = us[n];
            u.x = fma(lds, s.x, u.x);
            u.y = fma(lds, s.y, u.y);
        }
        const double2 a = p[n], c = q[n];
        const double tx = a.x * u.x - a.y * u.y, ty = a.x * u.y + a.y * u.x;  // x_p U1
        ax += tx * c.x - ty * c.y;
        ay += tx * c.y + ty * c.x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
    }
    if (lane == 0) out[(size_t)b * out_sample_stride + (size_t)j * out_j_stride] = make_double2(ax * scale, ay * scale);
}

// dsigma/dOmega(theta) = 10 xs_factor sum_{m,m'} | sum_{l < lmax, j} G[m][m'][l][j][theta] T[l][j] |^2   (:375-392)
// one thread per (sample, angle); G is read coalesced over the angles, T as broadcasts
__global__ void __launch_bounds__(128) qe_xs_kernel(int lmax, int nth, long long batch, const double2* __restrict__ T,
                                                    const double2* __restrict__ G, double factor, double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= batch * nth) return;
    const long long b = idx / nth;
    const int th = (int)(idx - b * nth);
    const int L1 = lmax + 1;
    const double2* t = T + (size_t)b * L1 * 2;
    double sum = 0.0;
    for (int mm = 0; mm < 4; ++mm) {
        double ax = 0.0, ay = 0.0;
        const double2* g = G + (size_t)mm * L1 * 2 * nth + th;
        for (int lj = 0; lj < 2 * lmax; ++lj) {  // l = 0 .. lmax - 1, as the reference's loop
            const double2 gv = g[(size_t)lj * nth], tv = t[lj];
            ax += gv.x * tv.x - gv.y * tv.y;
            ay += gv.x * tv.y + gv.y * tv.x;
        }
        sum += ax * ax + ay * ay;
    }
    out[idx] = factor * sum;
}

}  // namespace jitr

using namespace jitr;

extern "C" int jitr_b200_qe_overlap(int nbasis, long long batch, int nj, const double* xp, const double* xn, const double* U1c,
                                    const double* U1so, double lds0, double lds1, double scale, double* out,
                                    long long out_sample_stride, long long out_j_stride, void* stream) {
    if (batch == 0) return JITR_B200_OK;
    if (nbasis < 1 || batch < 0 || nj < 1 || nj > 2 || !xp || !xn || !U1c || !out)
        return set_error(JITR_B200_ERR_BAD_ARG, "qe_overlap: null pointer or bad size (nj = 1 or 2)");
    if (batch == 0) return JITR_B200_OK;
    const long long warps = batch * nj;
    const unsigned blocks = (unsigned)((warps + kQeWarps - 1) / kQeWarps);
    qe_overlap_kernel<<<blocks, kQeWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        nbasis, batch, nj, reinterpret_cast<const double2*>(xp), reinterpret_cast<const double2*>(xn),
        reinterpret_cast<const double2*>(U1c), reinterpret_cast<const double2*>(U1so), lds0, lds1, scale,
        reinterpret_cast<double2*>(out), out_sample_stride, out_j_stride);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("qe_overlap launch");
    return JITR_B200_OK;
}

extern "C" int jitr_b200_qe_xs(int lmax, int n_angles, long long batch, const double* T, const double* G, double factor,
                               double* out, void* stream) {
    if (batch == 0) return JITR_B200_OK;
    if (lmax < 0 || n_angles < 1 || batch < 0 || !T || !G || !out)
        return set_error(JITR_B200_ERR_BAD_ARG, "qe_xs: null pointer or bad size");
    if (batch == 0) return JITR_B200_OK;
    const long long n = batch * n_angles;
    qe_xs_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        lmax, n_angles, batch, reinterpret_cast<const double2*>(T), reinterpret_cast<const double2*>(G), factor, out);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("qe_xs launch");
    return JITR_B200_OK;
}
