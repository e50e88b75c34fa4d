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

struct HelmParams {
    double s_df, s_cf;   // 1/l_df^2, 1/l_cf^2
    double w_df, w_cf;   // ratio/l_df^2, (1-ratio)/l_cf^2
    double l_df, l_cf, ratio;
    int same_len;        // l_df == l_cf: one exp per pair
};

inline HelmParams make_helm(double l_df, double l_cf, double ratio) {
    HelmParams p;
    p.l_df = l_df; p.l_cf = l_cf; p.ratio = ratio;
    p.s_df = 1.0 / (l_df * l_df);
    p.s_cf = 1.0 / (l_cf * l_cf);
    p.w_df = ratio * p.s_df;
    p.w_cf = (1.0 - ratio) * p.s_cf;
    p.same_len = (l_df == l_cf);
    return p;
}

// exp(x) for x <= 0 (covariance envelopes).  Cody-Waite range reduction to
// |f| <= ln2/2 and a degree-13 Taylor/Horner polynomial: ~1 ulp, no table, no branches
// beyond the underflow clamp.  Cheaper than libdevice exp() on the shared FP64 pipe.
__device__ __forceinline__ double exp_neg(double x) {
    x = fmax(x, -745.0);
    const double L2E = 1.4426950408889634074;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    double t = rint(x * L2E);
    double f = fma(-t, LN2_HI, x);
    f = fma(-t, LN2_LO, f);
    double p = 1.605904383682161459939e-10;            // 1/13!
    p = fma(p, f, 2.087675698786809897921e-09);         // 1/12!
    p = fma(p, f, 2.505210838544171877505e-08);         // 1/11!
    p = fma(p, f, 2.755731922398589065256e-07);         // 1/10!
    p = fma(p, f, 2.755731922398589065256e-06);         // 1/9!
    p = fma(p, f, 2.480158730158730158730e-05);         // 1/8!
    p = fma(p, f, 1.984126984126984126984e-04);         // 1/7!
    p = fma(p, f, 1.388888888888888888889e-03);         // 1/6!
    p = fma(p, f, 8.333333333333333333333e-03);         // 1/5!
    p = fma(p, f, 4.166666666666666666667e-02);         // 1/4!
    p = fma(p, f, 1.666666666666666666667e-01);         // 1/3!
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    // scale by 2^t; t in [-1075, 0]: split so denormal results stay correct
    int ti = (int)t;
    int t1 = ti >> 1, t2 = ti - t1;
    double s1 = __longlong_as_double((long long)(1023 + t1) << 52);
    double s2 = __longlong_as_double((long long)(1023 + t2) << 52);
    return p * s1 * s2;
}

// 2x2 block for separation (d1, d2).
__device__ __forceinline__ void helm_block(const HelmParams& p, double d1, double d2,
                                           double& k11, double& k12, double& k22) {
    double a = d1 * d1, b = d2 * d2, c = d1 * d2;
    double r2 = a + b;
    double E = exp_neg(-0.5 * p.s_df * r2);
    double F = p.same_len ? E : exp_neg(-0.5 * p.s_cf * r2);
    double e = p.w_df * E, f = p.w_cf * F;
    double es = e * p.s_df, fs = f * p.s_cf;
    k11 = fma(-b, es, e) + fma(-a, fs, f);
    k22 = fma(-a, es, e) + fma(-b, fs, f);
    k12 = c * (es - fs);
}

// Derivatives of the block w.r.t. (l_df, l_cf, ratio).
//   compat == 0: analytic derivative
//       dK_df/dl = ratio e^{-C/2} [ (C-2)/l^3 A_df + (2/l^5)(r^2 I - B) ]
//       dK_cf/dl = (1-ratio) e^{-C/2} [ (C-2)/l^3 A_cf + (2/l^5) B ]
//   compat == 1: the integrands of the reference's update_gradients_full
//       (myKernel.py:77-81, 91-96): (2/l^3) G + A (2-C)/l^3
//   d/dratio = K_df - K_cf (myKernel.py:99-102) in both modes.
// g[p][0..2] = (11, 12, 22) entries for parameter p.
__device__ __forceinline__ void helm_block_grad(const HelmParams& p, int compat, double d1,
                                                double d2, double (&g)[3][3]) {
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
}

}  // namespace gp2d
