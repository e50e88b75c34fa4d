// Sum of space-time Helmholtz terms: the kernels krig.kriging(kernelType = 2, 3, 4, nKernels) asks
// its (absent) module myKernel2 for (krig.py:396-407):
//     k2 = myKernel2.divFreeK(input_dim=3, active_dims=[0,1,2], var=1., lt=1., ly=1., lx=1.)
//     k2 = myKernel2.curlFreeK(...)            k2 = divFreeK(...) + curlFreeK(...)
//     k  = k2 + k2 + ...                        (nKernels copies, independent parameters)
// Each term q has a type (divergence-free / curl-free), a variance and three length scales over
// the inputs (t, a, b); with s(d) = exp(-dt^2/2lt^2 - da^2/2la^2 - db^2/2lb^2) the 2x2 block of a
// term is what the stream function / potential construction of myKernel.py:39-52 gives for an
// anisotropic squared exponential:
//   div-free : var s [[(1 - db^2/lb^2)/lb^2,  da db/(la^2 lb^2)], [ . , (1 - da^2/la^2)/la^2]]
//   curl-free: var s [[(1 - da^2/la^2)/la^2, -da db/(la^2 lb^2)], [ . , (1 - db^2/lb^2)/lb^2]]
// The first component belongs to the first spatial coordinate, as everywhere in this library.
// With la == lb and no time factor, {div-free with var = ratio, curl-free with var = 1 - ratio}
// is exactly myKernel.myKernel.K (myKernel.py:27-53), which is how the family is pinned to the
// reference (tests/test_hsum_family.py); the anisotropic and time-dependent part is specified
// here because the reference's module is missing (parity unpinned for it).
#pragma once
#include "helmholtz.cuh"

namespace gp2d {

constexpr int HSUM_MAXQ = 8;
constexpr int HSUM_NP = 4;        // parameters per term: var, lt, la, lb

struct HsumTerm {
    double var, th, a1, a2;       // th = 1/(2 lt^2) (0 without time), a1 = 1/la^2, a2 = 1/lb^2
    double sgn;                   // +1 divergence-free, -1 curl-free
};

struct HsumParams {
    int Q, has_t, ldx, xo;
    HsumTerm t[HSUM_MAXQ];
    double kss0, kss1;            // prior variance of the two components (Kdiag)
};

// type[q]: 0 div-free, 1 curl-free; params[q] = (var, lt, la, lb); ldx = 2: points (a, b), lt ignored;
// ldx = 3: points (t, a, b)
inline bool make_hsum(int ldx, int Q, const int* type, const double* params, HsumParams* out) {
    if ((ldx != 2 && ldx != 3) || Q < 1 || Q > HSUM_MAXQ || !type || !params) return false;
    HsumParams p;
    p.Q = Q; p.has_t = ldx == 3; p.ldx = ldx; p.xo = ldx - 2;
    p.kss0 = p.kss1 = 0.0;
    for (int q = 0; q < HSUM_MAXQ; ++q) {
        HsumTerm& t = p.t[q];
        t.var = 0.0; t.th = 0.0; t.a1 = t.a2 = 1.0; t.sgn = 1.0;
        if (q >= Q) continue;
        const double var = params[4 * q], lt = params[4 * q + 1], la = params[4 * q + 2], lb = params[4 * q + 3];
        if (type[q] != 0 && type[q] != 1) return false;
        if (!(var >= 0.0) || !(la > 0.0) || !(lb > 0.0) || !isfinite(var) || !isfinite(la) || !isfinite(lb)) return false;
        if (p.has_t && (!(lt > 0.0) || !isfinite(lt))) return false;
        t.var = var;
        t.th = p.has_t ? 0.5 / (lt * lt) : 0.0;
        t.a1 = 1.0 / (la * la);
        t.a2 = 1.0 / (lb * lb);
        t.sgn = type[q] ? -1.0 : 1.0;
        p.kss0 += var * (type[q] ? t.a1 : t.a2);
        p.kss1 += var * (type[q] ? t.a2 : t.a1);
    }
    *out = p;
    return true;
}

// one term, separations (dt, d1, d2): block (k11, k12, k22) ADDED to the accumulators
__device__ __forceinline__ void hsum_term_add(const HsumTerm& t, double dt2, double a, double b, double c,
                                              double& k11, double& k12, double& k22,
                                              const double* __restrict__ tab = EXP2_TAB) {
    const double s1 = a * t.a1, s2 = b * t.a2;
    const double k = t.var * exp_neg(-fma(t.th, dt2, 0.5 * (s1 + s2)), tab);
    const double P = k * fma(-s1, t.a1, t.a1);        // k (1 - s1) / la^2
    const double R = k * fma(-s2, t.a2, t.a2);        // k (1 - s2) / lb^2
    const bool df = t.sgn > 0.0;
    k11 += df ? R : P;
    k22 += df ? P : R;
    k12 = fma(t.sgn * k, c * (t.a1 * t.a2), k12);
}

__device__ __forceinline__ HelmPoint hsum_point(const HsumParams& p, const double* __restrict__ X, long i) {
    const double* r = X + p.ldx * i;
    HelmPoint q;
    q.a = __ldg(r + p.xo);
    q.b = __ldg(r + p.xo + 1);
    q.t = p.has_t ? __ldg(r) : 0.0;
    return q;
}

__device__ __forceinline__ void hsum_block_pts(const HsumParams& p, const HelmPoint& x, const HelmPoint& y,
                                               double& k11, double& k12, double& k22,
                                               const double* __restrict__ tab = EXP2_TAB) {
    const double d1 = x.a - y.a, d2 = x.b - y.b, dt = x.t - y.t;
    const double a = d1 * d1, b = d2 * d2, c = d1 * d2, dt2 = dt * dt;
    k11 = k12 = k22 = 0.0;
    for (int q = 0; q < p.Q; ++q) hsum_term_add(p.t[q], dt2, a, b, c, k11, k12, k22, tab);
}

// Derivatives of one term's block w.r.t. (var, lt, la, lb), contracted with the weights
// (w11, w12 + w21, w22): acc[0..3] += sum over the block of dK/dparam * W.
__device__ __forceinline__ void hsum_term_grad(const HsumTerm& t, double dt2, double a, double b, double c,
                                               double w11, double ws, double w22, double (&acc)[HSUM_NP]) {
    const double s1 = a * t.a1, s2 = b * t.a2;
    const double E = exp_neg(-fma(t.th, dt2, 0.5 * (s1 + s2)));
    const double k = t.var * E;
    const double P = fma(-s1, t.a1, t.a1), R = fma(-s2, t.a2, t.a2), C = t.sgn * c * (t.a1 * t.a2);
    const bool df = t.sgn > 0.0;
    // weights seen by P, R, C
    const double wP = df ? w22 : w11, wR = df ? w11 : w22;
    const double blk = P * wP + R * wR + C * ws;              // block / k, contracted
    acc[0] += E * blk;                                        // d/dvar
    const double tt = 2.0 * t.th;
    acc[1] += k * blk * dt2 * tt * sqrt(tt);                  // d/dlt: block * dt^2 / lt^3
    const double i1 = sqrt(t.a1), i2 = sqrt(t.a2);            // 1/la, 1/lb
    // d(kP)/dla = k/la [s1 P + (4 s1 - 2)/la^2], d(kR)/dla = k/la s1 R, d(kC)/dla = k/la (s1 - 2) C
    acc[2] += k * i1 * (fma(s1, P, t.a1 * fma(4.0, s1, -2.0)) * wP + s1 * R * wR + (s1 - 2.0) * C * ws);
    acc[3] += k * i2 * (s2 * P * wP + fma(s2, R, t.a2 * fma(4.0, s2, -2.0)) * wR + (s2 - 2.0) * C * ws);
}

}  // namespace gp2d
