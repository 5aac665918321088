// Perey-Buck-type nonlocal kernels generated on the device, so that a nonlocal sweep is fed from a few numbers per sample
// instead of host-built N x N arrays.  Overflow-free restatement of perey_buck_nonlocal
// (optical_potentials/potential_forms.py:14-19:  exp(-(r^2 + r'^2)/beta^2) * 2 i^l z j_l(-i z) / (beta sqrt(pi)), whose
// spherical_jn of imaginary argument overflows beyond z ~ 700):  2 i^l z j_l(-i z) = 2 z i_l(z)  with the modified spherical
// Bessel function of the first kind, evaluated in its exponentially scaled form  e^{-z} i_l(z)  for ALL l = 0..lmax at once:
//   z >= 1/2 : Miller's downward recurrence  f_{l-1} = f_{l+1} + (2l+1)/z f_l  from  l = lmax + sqrt(80 z) + 20,
//              normalised with  e^{-z} i_0(z) = (1 - e^{-2z}) / (2z);
//   z <  1/2 : the ascending series  i_l(z) = z^l/(2l+1)!! sum_k (z^2/2)^k / (k! (2l+3)(2l+5)...(2l+2k+1)).
// The kernel of a sample is
//   U_l(r, r') = -(V + iW) f((r + r')/2; R, a) * 2 z e^{-z} i_l(z) exp(z - (r^2 + r'^2)/beta^2) / (beta sqrt(pi)),
//   z = zfac r r'/beta^2   (zfac = 2: the Gaussian nonlocality, exponent -(r - r')^2/beta^2;  zfac = 2 pi: the reference's literal form),
// f the Woods-Saxon form factor of potential_forms.py:40-52 (a <= 0: f = 1).
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"
#include "jitr_b200.h"

namespace jitr {
namespace {

constexpr int kGenMaxL = 64;

// out[l] = e^{-z} i_l(z), l = 0..lmax
__device__ __forceinline__ void scaled_bessel_i(const double z, const int lmax, double* __restrict__ out) {
    if (z < 0.5) {
        const double h = 0.5 * z * z;
        double pref = 1.0, ez = exp(-z);
        for (int l = 0; l <= lmax; ++l) {
            if (l > 0) pref *= z / (2 * l + 1);
            double term = 1.0, sum = 1.0;
            for (int k = 1; k < 24; ++k) {
                term *= h / (k * (double)(2 * l + 2 * k + 1));
                sum += term;
                if (term < 1e-17 * sum) break;
            }
            out[l] = pref * sum * ez;
        }
        return;
    }
    const int lstart = lmax + (int)sqrt(80.0 * z) + 20;
    double fp = 0.0, f = 1e-250;
    const double zi = 1.0 / z;
    for (int l = lstart; l > lmax; --l) {  // f = f_l, fp = f_{l+1}  ->  f_{l-1}
        const double fm = fma((double)(2 * l + 1) * zi, f, fp);
        fp = f;
        f = fm;
    }
    out[lmax] = f;
    for (int l = lmax; l > 0; --l) {
        const double fm = fma((double)(2 * l + 1) * zi, f, fp);
        fp = f;
        f = fm;
        out[l - 1] = f;
    }
    const double f0 = -expm1(-2.0 * z) / (2.0 * z);
    const double s = f0 / f;
    for (int l = 0; l <= lmax; ++l) out[l] *= s;
}

__global__ void perey_buck_kernel(const int N, const int lmax, const long long n_samples, const double* __restrict__ rgrid,
                                  const double* __restrict__ params, const double zfac, double2* __restrict__ out,
                                  const long long sample_stride, const long long l_stride) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)N * N;
    if (idx >= n_samples * per) return;
    const long long s = idx / per;
    const int nm = (int)(idx - s * per);
    const int n = nm / N, m = nm - n * N;
    const double* p = params + s * 5;
    const double V = p[0], W = p[1], beta = p[2], R = p[3], a = p[4];
    const double r = rgrid[n], rp = rgrid[m];
    const double b2 = beta * beta;
    const double z = zfac * r * rp / b2;
    // exp(z - (r^2 + r'^2)/beta^2); for zfac = 2 the exponent is -(r - r')^2 / beta^2 exactly
    const double ex = (zfac == 2.0) ? -((r - rp) * (r - rp)) / b2 : z - (r * r + rp * rp) / b2;
    double ws = 1.0;
    if (a > 0.0) {
        const double x = (0.5 * (r + rp) - R) / a;
        ws = x <= 36.841361487904734 ? 1.0 / (1.0 + exp(x)) : 0.0;  // MAX_ARG = ln 1e16 (potential_forms.py:11,40-52)
    }
    const double amp = ws * 2.0 * z * exp(ex) / (beta * 1.7724538509055159);
    double f[kGenMaxL + 1];
    if (amp != 0.0) scaled_bessel_i(z, lmax, f);
    double2* o = out + s * sample_stride + nm;
    for (int l = 0; l <= lmax; ++l) {
        const double h = amp != 0.0 ? amp * f[l] : 0.0;
        o[(long long)l * l_stride] = make_double2(-V * h, -W * h);
    }
}

}  // namespace
}  // namespace jitr

using namespace jitr;

extern "C" int jitr_b200_perey_buck_kernels(int nbasis, int lmax, long long n_samples, const double* rgrid, const double* params,
                                            double zfac, double* out, long long sample_stride, long long l_stride, void* stream) {
    if (n_samples == 0) return JITR_B200_OK;
    if (nbasis < 1 || lmax < 0 || lmax > kGenMaxL || n_samples < 0 || !rgrid || !params || !out || !(zfac > 0.0))
        return set_error(JITR_B200_ERR_BAD_ARG, "perey_buck_kernels: null pointer or bad size (lmax <= 64)");
    const long long total = n_samples * nbasis * nbasis;
    perey_buck_kernel<<<(unsigned)((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        nbasis, lmax, n_samples, rgrid, params, zfac, reinterpret_cast<double2*>(out),
        sample_stride > 0 ? sample_stride : (long long)(lmax + 1) * nbasis * nbasis, l_stride > 0 ? l_stride : (long long)nbasis * nbasis);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("perey_buck_kernels launch");
    return JITR_B200_OK;
}
