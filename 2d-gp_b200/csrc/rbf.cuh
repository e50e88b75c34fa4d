// Scalar kernel family of the reference's production predict path: a sum of Q anisotropic
// ("ARD") squared-exponential kernels over D input dimensions (t, y, x in krig),
//   k(x, x') = sum_q var_q * exp(-1/2 sum_d ((x_d - x'_d) / l_{q,d})^2)
// = GPy.kern.RBF(input_dim=3, ARD=True) summed nKernels times (krig.py:388,405-407) and
// scikit-learn's HP[0]*RBF(length_scale=[..]) + HP[4]*RBF([..]) (krig.py:174-178).  The white
// noise term (WhiteKernel, krig.py:179; GPy Gaussian_noise) is the diagonal shift of the fit.
#pragma once
#include "helmholtz.cuh"

namespace gp2d {

constexpr int RBF_MAXQ = 4, RBF_MAXD = 4;

struct RbfParams {
    int Q, D;
    double var[RBF_MAXQ];
    double ls[RBF_MAXQ][RBF_MAXD];
    double inv[RBF_MAXQ][RBF_MAXD];     // 1 / l
    double kss;                         // k(x, x) = sum_q var_q
};

inline bool make_rbf(int D, int Q, const double* var, const double* ls, RbfParams* out) {
    if (D < 1 || D > RBF_MAXD || Q < 1 || Q > RBF_MAXQ || !var || !ls) return false;
    RbfParams p;
    p.Q = Q; p.D = D; p.kss = 0.0;
    for (int q = 0; q < RBF_MAXQ; ++q) {
        p.var[q] = q < Q ? var[q] : 0.0;
        if (q < Q && !(var[q] >= 0.0)) return false;
        if (q < Q) p.kss += var[q];
        for (int d = 0; d < RBF_MAXD; ++d) {
            const bool on = q < Q && d < D;
            const double l = on ? ls[q * D + d] : 1.0;
            if (on && !(l > 0.0)) return false;
            p.ls[q][d] = l;
            p.inv[q][d] = on ? 1.0 / l : 0.0;
        }
    }
    *out = p;
    return true;
}

// a, b: coordinates padded with zeros to RBF_MAXD
__device__ __forceinline__ double rbf_eval(const RbfParams& p, const double (&a)[RBF_MAXD], const double (&b)[RBF_MAXD],
                                           const double* __restrict__ tab = EXP2_TAB) {
    double k = 0.0;
#pragma unroll
    for (int q = 0; q < RBF_MAXQ; ++q) {
        if (q < p.Q) {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < RBF_MAXD; ++d) {
                const double t = (a[d] - b[d]) * p.inv[q][d];
                s = fma(t, t, s);
            }
            k = fma(p.var[q], exp_neg(-0.5 * s, tab), k);
        }
    }
    return k;
}

__device__ __forceinline__ void rbf_load_point(const double* __restrict__ X, long i, int D, double (&x)[RBF_MAXD]) {
#pragma unroll
    for (int d = 0; d < RBF_MAXD; ++d) x[d] = d < D ? __ldg(X + i * D + d) : 0.0;      // read-only path, see helm_point
}

// gradient accumulators: per component q, slot q (1 + RBF_MAXD) is d/dvar_q and the next RBF_MAXD
// slots are d/dl_{q,d}:   dK/dvar_q = k_q / var_q ;  dK/dl_{q,d} = k_q delta_d^2 / l_{q,d}^3   (GPy RBF, ARD)
constexpr int RBF_NG = RBF_MAXQ * (1 + RBF_MAXD);

__device__ __forceinline__ void rbf_grad_terms(const RbfParams& p, const double (&a)[RBF_MAXD],
                                               const double (&b)[RBF_MAXD], double w, double (&acc)[RBF_NG]) {
#pragma unroll
    for (int q = 0; q < RBF_MAXQ; ++q) {
        if (q < p.Q) {
            double t2[RBF_MAXD], s = 0.0;
#pragma unroll
            for (int d = 0; d < RBF_MAXD; ++d) {
                const double t = (a[d] - b[d]) * p.inv[q][d];
                t2[d] = t * t;
                s += t2[d];
            }
            const double e = exp_neg(-0.5 * s) * w;
            acc[q * (1 + RBF_MAXD)] += e;                       // d/dvar_q
            const double ke = e * p.var[q];
#pragma unroll
            for (int d = 0; d < RBF_MAXD; ++d) acc[q * (1 + RBF_MAXD) + 1 + d] += ke * t2[d] * p.inv[q][d];   // t^2 / l
        }
    }
}


}  // namespace gp2d
