// Helmholtz (divergence-free + curl-free squared-exponential) 2x2 covariance block.
//
// Values follow the reference kernel myKernel.myKernel.K (myKernel.py:27-53):
//   K_df = e^{-r^2/2l_df^2}/l_df^2 [[1 - d2^2/l_df^2,  d1 d2/l_df^2], [ d1 d2/l_df^2, 1 - d1^2/l_df^2]]
//   K_cf = e^{-r^2/2l_cf^2}/l_cf^2 [[1 - d1^2/l_cf^2, -d1 d2/l_cf^2], [-d1 d2/l_cf^2, 1 - d2^2/l_cf^2]]
//   K    = ratio K_df + (1 - ratio) K_cf
// The block is symmetric (k12 == k21) and even in d, so K(a,b) == K(b,a).
#pragma once
#include <math.h>

namespace gp2d {

// Optional time factor (space-time product kernel, scratch.py:506-508: Kt(t) * nonDivK(y, x) with
// Kt = var exp(-dt^2 / 2 l_t^2) tiled over the 2x2 blocks, myKernel.py:350-360): points then have
// three columns (t, a, b) and every 2x2 block is multiplied by tau(dt) = tvar exp(-dt^2 / 2 lt^2).
struct HelmParams {
    double s_df, s_cf;   // 1/l_df^2, 1/l_cf^2
    double w_df, w_cf;   // ratio/l_df^2, (1-ratio)/l_cf^2
    double l_df, l_cf, ratio;
    int same_len;        // l_df == l_cf: one exp per pair
    int has_t;           // time factor on
    int ldx, xo;         // row stride of the point arrays and column of the first spatial coordinate
    double tvar, lt, thalf;   // thalf = 1 / (2 lt^2)
};

inline HelmParams make_helm(double l_df, double l_cf, double ratio) {
    HelmParams p;
    p.l_df = l_df; p.l_cf = l_cf; p.ratio = ratio;
    p.s_df = 1.0 / (l_df * l_df);
    p.s_cf = 1.0 / (l_cf * l_cf);
    p.w_df = ratio * p.s_df;
    p.w_cf = (1.0 - ratio) * p.s_cf;
    p.same_len = (l_df == l_cf);
    p.has_t = 0; p.ldx = 2; p.xo = 0;
    p.tvar = 1.0; p.lt = 1.0; p.thalf = 0.0;
    return p;
}

inline HelmParams make_helm_st(double l_df, double l_cf, double ratio, double tvar, double lt) {
    HelmParams p = make_helm(l_df, l_cf, ratio);
    p.has_t = 1; p.ldx = 3; p.xo = 1;
    p.tvar = tvar; p.lt = lt; p.thalf = 0.5 / (lt * lt);
    return p;
}

// exp(x) for x <= 0 (covariance envelopes): x = (64 k + j) ln2/64 + r, |r| <= ln2/128,
//   exp(x) = 2^k * 2^(j/64) * p5(r)
// 64-entry table of correctly rounded 2^(j/64), two-term Cody-Waite reduction and a
// degree-5 polynomial (truncation 3.5e-17): ~1 ulp with 10 FP64-pipe operations and a
// dependency chain half as long as a table-free degree-13 evaluation.  The DMMA and
// DFMA instructions share one FP64 pipe on sm_100a (measured), so every operation
// saved here is tensor-pipe time in the fused predictive kernel.
__device__ const double EXP2_TAB[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

// tab: the 64-entry table above, or a copy of it in shared memory (the fused predictive kernel
// leaves the L1 only ~30 KB next to its 221 KB of shared memory, and its streaming panel stores
// evict the table: every lookup then pays an L2 round trip on the critical path)
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ tab = EXP2_TAB) {
    x = fmax(x, -708.0);
    const double INV = 0x1.71547652b82fep+6;      // 64 / ln 2
    const double C_HI = 0x1.62e42fef80000p-7;    // ln2/64, low 18 bits zero: n * C_HI exact
    const double C_LO = 0x1.1cf79abc9e3b4p-42;
    const double nd = rint(x * INV);
    const int ni = (int)nd;
    double r = fma(-nd, C_HI, x);
    r = fma(-nd, C_LO, r);
    const double r2 = r * r;
    const double a = fma(r, 1.0 / 6.0, 0.5);            // 1/2 + r/6
    const double b = fma(r, 1.0 / 120.0, 1.0 / 24.0);   // 1/24 + r/120
    double p = fma(r2, fma(r2, b, a), r) + 1.0;         // 1 + r + r^2 a + r^4 b
    // 2^(j/64) with k added to its exponent field (k >= -1022: stays normal)
    const long long tb = __double_as_longlong(tab[ni & 63]) + ((long long)(ni >> 6) << 52);
    return p * __longlong_as_double(tb);
}

// 2x2 block for separation (d1, d2).
__device__ __forceinline__ void helm_block(const HelmParams& p, double d1, double d2,
                                           double& k11, double& k12, double& k22,
                                           const double* __restrict__ tab = EXP2_TAB) {
    double a = d1 * d1, b = d2 * d2, c = d1 * d2;
    double r2 = a + b;
    double E = exp_neg(-0.5 * p.s_df * r2, tab);
    double F = p.same_len ? E : exp_neg(-0.5 * p.s_cf * r2, tab);
    double e = p.w_df * E, f = p.w_cf * F;
    double es = e * p.s_df, fs = f * p.s_cf;
    k11 = fma(-b, es, e) + fma(-a, fs, f);
    k22 = fma(-a, es, e) + fma(-b, fs, f);
    k12 = c * (es - fs);
}

// A point of either layout: [N,2] rows (a, b) or [N,3] rows (t, a, b).
struct HelmPoint { double a, b, t; };
__device__ __forceinline__ HelmPoint helm_point(const HelmParams& p, const double* __restrict__ X, long i) {
    // read-only path (ld.global.nc): point arrays are never written by the kernels that read them,
    // and the compiler may then hoist these loads over unrelated stores (unrolled panel generation)
    const double* r = X + p.ldx * i;
    HelmPoint q;
    q.a = __ldg(r + p.xo);
    q.b = __ldg(r + p.xo + 1);
    q.t = p.has_t ? __ldg(r) : 0.0;
    return q;
}
__device__ __forceinline__ double helm_tau(const HelmParams& p, double dt, const double* __restrict__ tab = EXP2_TAB) {
    return p.tvar * exp_neg(-p.thalf * dt * dt, tab);
}
// 2x2 block between two points, time factor included when it is on.
__device__ __forceinline__ void helm_block_pts(const HelmParams& p, const HelmPoint& x, const HelmPoint& y,
                                               double& k11, double& k12, double& k22,
                                               const double* __restrict__ tab = EXP2_TAB) {
    helm_block(p, x.a - y.a, x.b - y.b, k11, k12, k22, tab);
    if (p.has_t) {
        const double tau = helm_tau(p, x.t - y.t, tab);
        k11 *= tau; k12 *= tau; k22 *= tau;
    }
}

// Derivatives of the block w.r.t. (l_df, l_cf, ratio).
//   compat == 0: analytic derivative
//       dK_df/dl = ratio e^{-C/2} [ (C-2)/l^3 A_df + (2/l^5)(r^2 I - B) ]
//       dK_cf/dl = (1-ratio) e^{-C/2} [ (C-2)/l^3 A_cf + (2/l^5) B ]
//   compat == 1: the integrands of the reference's update_gradients_full
//       (myKernel.py:77-81, 91-96): (2/l^3) G + A (2-C)/l^3
//   d/dratio = K_df - K_cf (myKernel.py:99-102) in both modes.
// g[p][0..2] = (11, 12, 22) entries for parameter p.
// g[3], g[4]: derivatives w.r.t. (tvar, lt) of the time factor (zero when it is off); with the
// factor on, g[0..2] carry tau as well.
constexpr int HELM_NP = 5;
__device__ __forceinline__ void helm_block_grad(const HelmParams& p, int compat, double d1,
                                                double d2, double dt, double (&g)[HELM_NP][3]) {
    double a = d1 * d1, b = d2 * d2, c = d1 * d2, r2 = a + b;
    {   // divergence-free part
        double l = p.l_df, s = p.s_df, l3 = l * l * l;
        double C = r2 * s, E = exp_neg(-0.5 * C);
        double A11 = 1.0 - b * s, A22 = 1.0 - a * s, A12 = c * s;   // B/l^2 + (1-C) I
        double G11 = b, G22 = a, G12 = -c;                         // r^2 I - B
        double ca, cg;
        if (compat) { ca = (2.0 - C) / l3; cg = 2.0 / l3; }
        else        { ca = (C - 2.0) / l3; cg = 2.0 / (l3 * l * l); }
        double w = p.ratio * E;
        g[0][0] = w * (ca * A11 + cg * G11);
        g[0][1] = w * (ca * A12 + cg * G12);
        g[0][2] = w * (ca * A22 + cg * G22);
        g[2][0] = s * E * A11; g[2][1] = s * E * A12; g[2][2] = s * E * A22;
    }
    {   // curl-free part
        double l = p.l_cf, s = p.s_cf, l3 = l * l * l;
        double C = r2 * s, E = exp_neg(-0.5 * C);
        double A11 = 1.0 - a * s, A22 = 1.0 - b * s, A12 = -c * s;  // I - B/l^2
        double ca, cg;
        if (compat) { ca = (2.0 - C) / l3; cg = 2.0 / l3; }
        else        { ca = (C - 2.0) / l3; cg = 2.0 / (l3 * l * l); }
        double w = (1.0 - p.ratio) * E;
        g[1][0] = w * (ca * A11 + cg * a);
        g[1][1] = w * (ca * A12 + cg * c);
        g[1][2] = w * (ca * A22 + cg * b);
        g[2][0] -= s * E * A11; g[2][1] -= s * E * A12; g[2][2] -= s * E * A22;
    }
#pragma unroll
    for (int e = 0; e < 3; ++e) g[3][e] = g[4][e] = 0.0;
    if (p.has_t) {
        // spatial block K_sp = ratio K_df + (1-ratio) K_cf; K = tau K_sp
        double k11, k12, k22;
        helm_block(p, d1, d2, k11, k12, k22);
        const double et = exp_neg(-p.thalf * dt * dt), tau = p.tvar * et;
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int e = 0; e < 3; ++e) g[q][e] *= tau;
        const double cl = tau * dt * dt / (p.lt * p.lt * p.lt);      // dtau/dlt = tau dt^2 / lt^3
        g[3][0] = et * k11; g[3][1] = et * k12; g[3][2] = et * k22;
        g[4][0] = cl * k11; g[4][1] = cl * k12; g[4][2] = cl * k22;
    }
}

}  // namespace gp2d
