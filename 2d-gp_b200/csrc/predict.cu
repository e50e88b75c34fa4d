// Fused predictive pass.  Persistent CTAs (one per SM: 8 consumer warps + 1 producer warp) take work
// items (column tile, split); a column tile is 64 grid points x 2 components = 128 stacked columns
// (128 grid points for the scalar family).  Two phases per item:
//   1. the consumer warps generate the K* panel [npad x 128] ONCE (1-2 exps per point pair give the
//      four entries of a 2x2 block thanks to the pair-interleaved internal ordering), already in the
//      swizzled shared-memory tile image, into the CTA's private scratch panel, and accumulate
//      mean = K* alpha on the way (observations staged through shared memory, branch-free pair loop);
//   2. V = Z K*^T on the DMMA pipe (Z = L^-1), row block after row block: one producer thread
//      streams 16 KB tiles of Zt and of the panel with bulk copies into a 6-stage mbarrier ring,
//      the consumers keep a 128 x 128 accumulator tile in registers and fold sum(v^2) per row
//      block into fixed row-block groups, so var = k** - colsumsq(V) needs no atomics and is
//      bit-reproducible for any grid size, partition or split.
// Replaces GPy model.predict (krig.py:543-544; GP_plots.py:768), GP_scripts.getMean +
// the diagonal of GP_laser.py:129-131, and sklearn predict(return_std=True)
// (krig.py:194; algebra of _gpr.py:446-491).
//
// Why a scratch panel: DMMA and DFMA share one FP64 pipe on sm_100a (measured: 37 TF/s
// each, 36 TF/s mixed).  Regenerating K* tiles inside the k-loop of every row block cost
// 2 x (n/256) exps per pair and 20% of the kernel; generating once costs ~1% and the
// re-reads (once per row block, from L2 / DRAM) use 15 % of the DRAM bandwidth.
// The full K* (n x 2M) still never exists: only one [npad x 128] panel per resident CTA.
#include "common.cuh"
#include "dgemm.cuh"
#include "linalg.h"
#include "pipeline.cuh"

namespace gp2d {

enum : int { FAM_HELM = 0, FAM_RBF = 1 };

struct PredictArgs {
    const double* Zt; int npad;     // L^-1 as pre-swizzled 128 x 16 tiles, row block major (linalg.h)
    const double* alpha;          // internal order (pair-interleaved for the vector kernel), zero padded to npad
    const double* X; int N;       // observation points [N,2] (Helmholtz) or [N,D] (RBF)
    HelmParams hp;
    RbfParams rp;
    HsumParams sp; int use_hsum;  // FAM_HELM tiles with the term-sum generator (hsum.cuh)
    const double* Xs; int M;
    long out_stride;              // component stride of mean/var
    double kss, kss1, var_add;    // prior variance of component 0 (and of the scalar family) / component 1
    double* mean; double* var;
    double* scratch;              // gridDim.x panels of npad x 128 doubles
    double* partial;              // [ntiles][8 groups][128] column sums (nsplit > 1)
    int ntiles, nsplit;
    const int* gate;              // non-null: run only when *gate == 0 (the int8 kernel of predict_i8.cu owns the other values)
};

// Row-block groups.  The column sums of V^2 are ALWAYS formed as sum_{g=0..7} (group g), where
// group g owns the row blocks {16 b + g, 16 b + 15 - g : b = 0, 1, ...} (pairs of equal total
// cost in the triangle).  A work item is (column tile, split s of nsplit in {1,2,4,8}) and covers
// the groups [s 8/nsplit, (s+1) 8/nsplit): with few column tiles (small grids) the row blocks of
// one tile are spread over several CTAs, and because the grouping and the order of the final
// sum never change, the result is bit-identical for every nsplit and every grid partition.
constexpr int PRED_GROUPS = 8;

// shared memory: operand ring | mbarriers | group column sums [8][2][128] | mean partials [4][64][2]
constexpr int PRED_MAX_ROWBLOCKS = 1024;          // npad <= 131072
constexpr int PRED_STAGE_OBS = 256;                // observations staged in shared memory at a time (phase 1)
constexpr int PRED_PARAM_DOUBLES = 64;             // shared-memory copy of HelmParams / RbfParams
constexpr int PRED_RED_DOUBLES = PRED_GROUPS * 2 * TILE + 4 * 64 * 2 + PRED_MAX_ROWBLOCKS / 2 + 5 * PRED_STAGE_OBS + PRED_PARAM_DOUBLES;
static_assert(sizeof(HelmParams) <= PRED_PARAM_DOUBLES * sizeof(double) && sizeof(RbfParams) <= PRED_PARAM_DOUBLES * sizeof(double) &&
              sizeof(HsumParams) <= PRED_PARAM_DOUBLES * sizeof(double), "parameter block");
constexpr int PRED_SMEM_BYTES = WS_RING_BYTES + WS_BAR_BYTES + PRED_RED_DOUBLES * (int)sizeof(double);

__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;\n" ::"n"(WS_CONSUMERS) : "memory"); }

// Phase 1 stages the observations through shared memory, 256 at a time: coordinates and alpha are
// loaded once, coalesced, by the 256 consumer threads and then read as broadcasts.  Reading them
// from global memory inside the pair loop put an L2 round trip at the head of every pair (the L1
// is ~30 KB next to 221 KB of shared memory and the streaming panel stores keep evicting it): 58 %
// of the phase-1 stall samples in the round-1 profile.
//
// The pair loops are branch-free (validity is a 0/1 factor, the kernel variant is a template
// parameter): with control flow in the body the compiler unrolls but does not interleave the
// iterations, and one dependent FP64 chain per thread took ~840 cycles per pair.
//
// Scalar family.  Thread (gjl, os): grid points gp0 + 2 gjl + {0,1} (one 16-byte chunk of every
// panel row), observations os, os + 4, ...
template <int Q>
__device__ __forceinline__ double rbf_eval_q(const RbfParams& p, const double (&a)[RBF_MAXD], const double (&b)[RBF_MAXD]) {
    double k = 0.0;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < RBF_MAXD; ++d) {
            const double t = (a[d] - b[d]) * p.inv[q][d];
            s = fma(t, t, s);
        }
        k = fma(p.var[q], exp_neg(-0.5 * s), k);
    }
    return k;
}

template <int Q>
__device__ __forceinline__ void rbf_phase1_loop(const PredictArgs& p, const RbfParams& rp, double* __restrict__ panel,
                                                double* __restrict__ stage, int gp0, int tid, double& mu0, double& mu1) {
    constexpr int OS = WS_CONSUMERS / 64;
    const int gjl = tid & 63, os = tid >> 6;
    const int gj = gp0 + 2 * gjl;
    const double w0 = gj < p.M ? 1.0 : 0.0, w1 = gj + 1 < p.M ? 1.0 : 0.0;
    double b0[RBF_MAXD], b1[RBF_MAXD];
    rbf_load_point(p.Xs, gj < p.M ? gj : 0, rp.D, b0);
    rbf_load_point(p.Xs, gj + 1 < p.M ? gj + 1 : 0, rp.D, b1);
    double m0 = 0.0, m1 = 0.0;
    for (int ob = 0; ob < p.npad; ob += PRED_STAGE_OBS) {
        consumer_barrier();                     // the previous batch has been read
        {
            const int o = ob + tid;
            double a[RBF_MAXD];
            rbf_load_point(p.X, o < p.N ? o : 0, rp.D, a);
#pragma unroll
            for (int d = 0; d < RBF_MAXD; ++d) stage[tid * 5 + d] = a[d];
            stage[tid * 5 + 4] = o < p.N ? __ldg(p.alpha + o) : 0.0;
        }
        consumer_barrier();
        const int nloc = min(PRED_STAGE_OBS, p.npad - ob);
#pragma unroll 4
        for (int ol = os; ol < nloc; ol += OS) {
            const int o = ob + ol;
            double a[RBF_MAXD];
#pragma unroll
            for (int d = 0; d < RBF_MAXD; ++d) a[d] = stage[ol * 5 + d];
            const double wo = o < p.N ? 1.0 : 0.0;
            const double k0 = rbf_eval_q<Q>(rp, a, b0) * (wo * w0);
            const double k1 = rbf_eval_q<Q>(rp, a, b1) * (wo * w1);
            const double al = stage[ol * 5 + 4];
            m0 = fma(k0, al, m0);
            m1 = fma(k1, al, m1);
            *reinterpret_cast<double2*>(panel + (size_t)o * TILE + 2 * (gjl ^ ((o & 3) << 1))) = make_double2(k0, k1);
        }
    }
    fence_proxy_async();
    mu0 = m0;
    mu1 = m1;
}

// out of line: the register needs of the generators must not push the accumulator loop into spilling
__device__ __noinline__ void rbf_phase1(const PredictArgs& p, const RbfParams& rp, double* __restrict__ panel,
                                        double* __restrict__ stage, int gp0, int tid, double& mu0, double& mu1) {
    switch (rp.Q) {
        case 1: rbf_phase1_loop<1>(p, rp, panel, stage, gp0, tid, mu0, mu1); break;
        case 2: rbf_phase1_loop<2>(p, rp, panel, stage, gp0, tid, mu0, mu1); break;
        case 3: rbf_phase1_loop<3>(p, rp, panel, stage, gp0, tid, mu0, mu1); break;
        default: rbf_phase1_loop<4>(p, rp, panel, stage, gp0, tid, mu0, mu1); break;
    }
}

// Helmholtz family.  Thread (gjl, os): grid point gp0 + gjl (chunk gjl of every panel row),
// observations os, os + 4, ...; rows 2o, 2o+1 of the panel.
template <bool SAME_LEN, bool HAS_T>
__device__ __forceinline__ void helm_phase1_loop(const PredictArgs& p, const HelmParams& hp, double* __restrict__ panel,
                                                 double* __restrict__ stage, int gp0, int tid, double& mu0, double& mu1) {
    constexpr int OS = WS_CONSUMERS / 64;
    const int gjl = tid & 63, os = tid >> 6;
    const int gj = gp0 + gjl;
    const double wg = gj < p.M ? 1.0 : 0.0;
    const HelmPoint gpt = helm_point(hp, p.Xs, gj < p.M ? gj : 0);
    const int nobs_pad = p.npad >> 1;
    double m0 = 0.0, m1 = 0.0;
    for (int ob = 0; ob < nobs_pad; ob += PRED_STAGE_OBS) {
        consumer_barrier();                     // the previous batch has been read
        {
            const int o = ob + tid;
            const bool ov = o < p.N;
            const HelmPoint q = helm_point(hp, p.X, ov ? o : 0);
            stage[tid * 5 + 0] = q.a;
            stage[tid * 5 + 1] = q.b;
            stage[tid * 5 + 2] = q.t;
            stage[tid * 5 + 3] = ov ? __ldg(p.alpha + 2 * o) : 0.0;
            stage[tid * 5 + 4] = ov ? __ldg(p.alpha + 2 * o + 1) : 0.0;
        }
        consumer_barrier();
        const int nloc = min(PRED_STAGE_OBS, nobs_pad - ob);
#pragma unroll 4
        for (int ol = os; ol < nloc; ol += OS) {
            const int o = ob + ol;
            const double d1 = stage[ol * 5 + 0] - gpt.a, d2 = stage[ol * 5 + 1] - gpt.b;
            // helm_block, specialised at compile time
            const double a = d1 * d1, b = d2 * d2, c = d1 * d2, r2 = a + b;
            const double E = exp_neg(-0.5 * hp.s_df * r2);
            const double F = SAME_LEN ? E : exp_neg(-0.5 * hp.s_cf * r2);
            double w = o < p.N ? wg : 0.0;
            if (HAS_T) {
                const double dt = stage[ol * 5 + 2] - gpt.t;
                w *= hp.tvar * exp_neg(-hp.thalf * dt * dt);
            }
            const double e = hp.w_df * E, f = hp.w_cf * F;
            const double es = e * hp.s_df, fs = f * hp.s_cf;
            double k11 = fma(-b, es, e) + fma(-a, fs, f);
            double k22 = fma(-a, es, e) + fma(-b, fs, f);
            double k12 = c * (es - fs);
            k11 *= w; k12 *= w; k22 *= w;
            const double a0 = stage[ol * 5 + 3], a1 = stage[ol * 5 + 4];
            m0 = fma(k11, a0, fma(k12, a1, m0));
            m1 = fma(k12, a0, fma(k22, a1, m1));
            // rows k = 2o, 2o+1 of the MN-major tile image: 16-byte chunk gjl, swizzled by k & 3
            double* r0 = panel + (size_t)(2 * o) * TILE;
            const int sw = (o & 1) << 2;
            *reinterpret_cast<double2*>(r0 + 2 * (gjl ^ sw)) = make_double2(k11, k12);
            *reinterpret_cast<double2*>(r0 + TILE + 2 * (gjl ^ (sw | 2))) = make_double2(k12, k22);
        }
    }
    fence_proxy_async();      // the panel is read back by bulk (async-proxy) copies
    mu0 = m0;
    mu1 = m1;
}

__device__ __noinline__ void helm_phase1(const PredictArgs& p, const HelmParams& hp, double* __restrict__ panel,
                                         double* __restrict__ stage, int gp0, int tid, double& mu0, double& mu1) {
    if (hp.has_t) {
        if (hp.same_len) helm_phase1_loop<true, true>(p, hp, panel, stage, gp0, tid, mu0, mu1);
        else helm_phase1_loop<false, true>(p, hp, panel, stage, gp0, tid, mu0, mu1);
    } else {
        if (hp.same_len) helm_phase1_loop<true, false>(p, hp, panel, stage, gp0, tid, mu0, mu1);
        else helm_phase1_loop<false, false>(p, hp, panel, stage, gp0, tid, mu0, mu1);
    }
}

// Sum of space-time Helmholtz terms: same thread mapping and panel image as helm_phase1_loop; one
// exponential per term and pair.  NQ > 0: compile-time term count (unrolled); NQ == 0: run-time loop.
template <int NQ, bool HAS_T>
__device__ __forceinline__ void hsum_phase1_loop(const PredictArgs& p, const HsumParams& sp, double* __restrict__ panel,
                                                 double* __restrict__ stage, int gp0, int tid, double& mu0, double& mu1) {
    constexpr int OS = WS_CONSUMERS / 64;
    const int gjl = tid & 63, os = tid >> 6;
    const int gj = gp0 + gjl;
    const double wg = gj < p.M ? 1.0 : 0.0;
    const HelmPoint gpt = hsum_point(sp, p.Xs, gj < p.M ? gj : 0);
    const int nobs_pad = p.npad >> 1;
    const int nq = NQ > 0 ? NQ : sp.Q;
    double m0 = 0.0, m1 = 0.0;
    for (int ob = 0; ob < nobs_pad; ob += PRED_STAGE_OBS) {
        consumer_barrier();                     // the previous batch has been read
        {
            const int o = ob + tid;
            const bool ov = o < p.N;
            const HelmPoint q = hsum_point(sp, p.X, ov ? o : 0);
            stage[tid * 5 + 0] = q.a;
            stage[tid * 5 + 1] = q.b;
            stage[tid * 5 + 2] = q.t;
            stage[tid * 5 + 3] = ov ? __ldg(p.alpha + 2 * o) : 0.0;
            stage[tid * 5 + 4] = ov ? __ldg(p.alpha + 2 * o + 1) : 0.0;
        }
        consumer_barrier();
        const int nloc = min(PRED_STAGE_OBS, nobs_pad - ob);
#pragma unroll 2
        for (int ol = os; ol < nloc; ol += OS) {
            const int o = ob + ol;
            const double d1 = stage[ol * 5 + 0] - gpt.a, d2 = stage[ol * 5 + 1] - gpt.b;
            const double a = d1 * d1, b = d2 * d2, c = d1 * d2;
            double dt2 = 0.0;
            if (HAS_T) {
                const double dt = stage[ol * 5 + 2] - gpt.t;
                dt2 = dt * dt;
            }
            double k11 = 0.0, k12 = 0.0, k22 = 0.0;
            if (NQ > 0) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) hsum_term_add(sp.t[q], dt2, a, b, c, k11, k12, k22);
            } else {
                for (int q = 0; q < nq; ++q) hsum_term_add(sp.t[q], dt2, a, b, c, k11, k12, k22);
            }
            const double w = o < p.N ? wg : 0.0;
            k11 *= w; k12 *= w; k22 *= w;
            const double a0 = stage[ol * 5 + 3], a1 = stage[ol * 5 + 4];
            m0 = fma(k11, a0, fma(k12, a1, m0));
            m1 = fma(k12, a0, fma(k22, a1, m1));
            double* r0 = panel + (size_t)(2 * o) * TILE;
            const int sw = (o & 1) << 2;
            *reinterpret_cast<double2*>(r0 + 2 * (gjl ^ sw)) = make_double2(k11, k12);
            *reinterpret_cast<double2*>(r0 + TILE + 2 * (gjl ^ (sw | 2))) = make_double2(k12, k22);
        }
    }
    fence_proxy_async();      // the panel is read back by bulk (async-proxy) copies
    mu0 = m0;
    mu1 = m1;
}

__device__ __noinline__ void hsum_phase1(const PredictArgs& p, const HsumParams& sp, double* __restrict__ panel,
                                         double* __restrict__ stage, int gp0, int tid, double& mu0, double& mu1) {
    if (sp.has_t) {
        switch (sp.Q) {
            case 1: hsum_phase1_loop<1, true>(p, sp, panel, stage, gp0, tid, mu0, mu1); break;
            case 2: hsum_phase1_loop<2, true>(p, sp, panel, stage, gp0, tid, mu0, mu1); break;
            default: hsum_phase1_loop<0, true>(p, sp, panel, stage, gp0, tid, mu0, mu1); break;
        }
    } else {
        switch (sp.Q) {
            case 1: hsum_phase1_loop<1, false>(p, sp, panel, stage, gp0, tid, mu0, mu1); break;
            case 2: hsum_phase1_loop<2, false>(p, sp, panel, stage, gp0, tid, mu0, mu1); break;
            default: hsum_phase1_loop<0, false>(p, sp, panel, stage, gp0, tid, mu0, mu1); break;
        }
    }
}

// 8 consumer warps (phase 1 + DMMA) and one producer warp (bulk-copy ring), see pipeline.cuh.
// FAM selects the covariance family of phase 1: the matrix-valued Helmholtz kernel (a column tile
// is 64 grid points x 2 components) or the scalar ARD-RBF sum (a column tile is 128 grid points).
template <int FAM>
__global__ void __launch_bounds__(WS_THREADS, 1) predict_kernel(const __grid_constant__ PredictArgs p) {
    constexpr int PTS = FAM == FAM_HELM ? 64 : 128;      // grid points per column tile
    constexpr int WM = 2, OS = WS_CONSUMERS / 64;
    if (p.gate && *p.gate != 0) return;
    extern __shared__ __align__(16) double smem[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + WS_STAGES * WS_STAGE_DOUBLES);
    WsBarriers wb{bars, bars + WS_STAGES};
    double* sh_grp = reinterpret_cast<double*>(bars + 2 * WS_STAGES);   // [8][WM][128]
    double* sh_mu = sh_grp + PRED_GROUPS * WM * TILE;                   // [OS][64][2]
    int* sh_rb = reinterpret_cast<int*>(sh_mu + OS * 64 * 2);           // row blocks of the item: li | group << 16
    double* sh_stage = reinterpret_cast<double*>(sh_rb + PRED_MAX_ROWBLOCKS);   // phase 1: 256 observations x 5 doubles
    double* sh_par = sh_stage + 5 * PRED_STAGE_OBS;                    // kernel parameters (a noinline callee would
                                                                       // otherwise read them through generic loads)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == WS_CONSUMERS / 32;
    // warps w and w+4 share a scheduler: give them different row halves (wm), so that when one
    // of them has nothing to do in a stage (see `skip` below) the other gets the whole FP64 pipe
    const int wm = (warp >> 2) & 1, wn = warp & 3;
    const int gjl = tid & 63, os = tid >> 6;
    const int nb = p.npad / TILE;
    const int gper = PRED_GROUPS / p.nsplit;
    double* panel = p.scratch + (size_t)blockIdx.x * ((size_t)p.npad * TILE + 20480 / sizeof(double));

    FragLane<false, 64> fa;
    FragLane<true, 32> fb;
    fa.init(wm, lane);
    fb.init(wn, lane);
    wb.init(tid, 1);
    {
        const bool hsum = FAM == FAM_HELM && p.use_hsum;
        const double* src = FAM == FAM_RBF ? reinterpret_cast<const double*>(&p.rp)
                            : hsum ? reinterpret_cast<const double*>(&p.sp) : reinterpret_cast<const double*>(&p.hp);
        const int nd = (int)((FAM == FAM_RBF ? sizeof(RbfParams) : hsum ? sizeof(HsumParams) : sizeof(HelmParams)) / sizeof(double));
        if (tid < nd) sh_par[tid] = src[tid];
    }
    __syncthreads();

    int rs = 0;            // ring cursor; both roles walk the same sequence of stages
    unsigned rph = 0;
    long fills = 0;

    for (int item = blockIdx.x; item < p.ntiles * p.nsplit; item += gridDim.x) {
        const int ct = item / p.nsplit, split = item - ct * p.nsplit;
        const int g0 = split * gper, g1 = g0 + gper;
        const int gp0 = ct * PTS;
        // ---------------- phase 1 (consumer warps): K* panel + mean ------------------------
        double mu0 = 0.0, mu1 = 0.0;
        if (!producer) {
            if (FAM == FAM_RBF) rbf_phase1(p, *reinterpret_cast<const RbfParams*>(sh_par), panel, sh_stage, gp0, tid, mu0, mu1);
            else if (p.use_hsum) hsum_phase1(p, *reinterpret_cast<const HsumParams*>(sh_par), panel, sh_stage, gp0, tid, mu0, mu1);
            else helm_phase1(p, *reinterpret_cast<const HelmParams*>(sh_par), panel, sh_stage, gp0, tid, mu0, mu1);
        }
        // the item's row blocks in processing order (both roles walk this list)
        int nrb = 0;
        for (int g = g0; g < g1; ++g)
            for (int base = 0; base < nb; base += 2 * PRED_GROUPS) {
                const int la = base + g, lb = base + 2 * PRED_GROUPS - 1 - g;
                if (la < nb) { if (tid == 0) sh_rb[nrb] = la | (g << 16); ++nrb; }
                if (lb < nb) { if (tid == 0) sh_rb[nrb] = lb | (g << 16); ++nrb; }
            }
        __syncthreads();      // panel (global) visible to the producer; previous item's reduce done
        if (!producer) {
            sh_mu[(os * 64 + gjl) * 2 + 0] = mu0;
            sh_mu[(os * 64 + gjl) * 2 + 1] = mu1;
        }

        if (producer) {
            // ---------------- phase 2, producer: one thread, two 16 KB bulk copies per stage ----
            if (lane == 0) {
                fence_proxy_async();
                for (int q = 0; q < nrb; ++q) {
                    const int li = sh_rb[q] & 0xffff;
                    // tiles of row block li are stored contiguously from tile 8 li (li+1)/2
                    const double* zt = p.Zt + (size_t)(4 * li * (li + 1)) * TILE_DOUBLES;
                    const int nkt = 8 * (li + 1);
                    for (int lkt = 0; lkt < nkt; ++lkt) {
                        if (fills >= WS_STAGES) mbar_wait(wb.empty + rs, rph ^ 1u);
                        double* st = smem + rs * WS_STAGE_DOUBLES;
                        mbar_arrive_expect_tx(wb.full + rs, 2 * TILE_DOUBLES * (unsigned)sizeof(double));
                        bulk_g2s(st, zt, TILE_DOUBLES * (unsigned)sizeof(double), wb.full + rs);
                        bulk_g2s(st + TILE_DOUBLES, panel + (size_t)lkt * TILE_DOUBLES,
                                 TILE_DOUBLES * (unsigned)sizeof(double), wb.full + rs);
                        zt += TILE_DOUBLES;
                        ++fills;
                        if (++rs == WS_STAGES) { rs = 0; rph ^= 1u; }
                    }
                }
            }
            __syncwarp();
        } else {
            // ---------------- phase 2, consumers: column sums of (Z K*^T)^2 -------------------
            double acc[8][4][2];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            // lane (lane>>2)==0 of warp (wm, wn) owns the 8 columns wn*32 + j*8 + 2*(lane&3) + e of
            // every group slot of its wm: it zeroes and accumulates them (no other writer)
            const int mycol = wm * TILE + wn * 32 + 2 * (lane & 3);
            if ((lane >> 2) == 0) {
                for (int g = 0; g < PRED_GROUPS; ++g)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        sh_grp[g * WM * TILE + mycol + j * 8] = 0.0;
                        sh_grp[g * WM * TILE + mycol + j * 8 + 1] = 0.0;
                    }
            }
            const int nvalid = FAM == FAM_HELM ? 2 * p.N : p.N;
            for (int q = 0; q < nrb; ++q) {
                    {
                        const int li = sh_rb[q] & 0xffff, g = sh_rb[q] >> 16;
                        const int nkt = 8 * (li + 1);
                        const int row0 = li * TILE + wm * 64;
                        // rows 32..63 of this warp tile are padding (Helmholtz instantiation only: the RBF one
                        // has no registers to spare for the second code path)
                        const bool half = FAM == FAM_HELM && nvalid - row0 <= 32;
                        for (int ckt = 0; ckt < nkt; ++ckt) {
                            mbar_wait(wb.full + rs, rph);
                            const double* st = smem + rs * WS_STAGE_DOUBLES;
                            // this warp's 64 rows of Z are exactly zero in this k-tile when the tile lies
                            // right of the diagonal (upper half of a diagonal block) and contribute nothing
                            // when they are identity padding (the matching panel rows are zero)
                            const bool skip = (ckt * BK > row0 + 63) || (row0 >= nvalid);
                            if (!skip) {
                                if (half) ws_mma_stage_half<false, true>(st, st + TILE_DOUBLES, fa, fb, acc);
                                else ws_mma_stage<false, true>(st, st + TILE_DOUBLES, fa, fb, acc);
                            }
                            __syncwarp();
                            if (lane == 0) mbar_arrive(wb.empty + rs);
                            if (++rs == WS_STAGES) { rs = 0; rph ^= 1u; }
                        }
                        // fold this row block: squares summed over the 8 blocks of the thread, then over
                        // the 8 row lanes (fixed order), then into the group slot
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                double v = 0.0;
#pragma unroll
                                for (int mb = 0; mb < 8; ++mb) {
                                    v = fma(acc[mb][j][e], acc[mb][j][e], v);
                                    acc[mb][j][e] = 0.0;
                                }
                                v += __shfl_xor_sync(0xffffffffu, v, 4);
                                v += __shfl_xor_sync(0xffffffffu, v, 8);
                                v += __shfl_xor_sync(0xffffffffu, v, 16);
                                if ((lane >> 2) == 0) sh_grp[g * WM * TILE + mycol + j * 8 + e] += v;
                            }
                    }
            }
        }
        __syncthreads();      // group sums visible; every bulk copy of this item has been consumed
        if (tid < TILE) {
            // column tid of the tile: Helmholtz (point tid>>1, component tid&1), RBF point tid
            const int pj = tid >> 1, c = tid & 1;
            const int j = FAM == FAM_HELM ? gp0 + pj : gp0 + tid;
            const long oidx = FAM == FAM_HELM ? (long)c * p.out_stride + j : (long)j;
            if (p.nsplit == 1) {
                if (j < p.M) {
                    double ss = 0.0;
                    for (int g = 0; g < PRED_GROUPS; ++g) ss += sh_grp[g * WM * TILE + tid] + sh_grp[g * WM * TILE + TILE + tid];
                    double v = (FAM == FAM_HELM && c ? p.kss1 : p.kss) - ss;
                    v = v < 0.0 ? 0.0 : v;
                    p.var[oidx] = v + p.var_add;
                }
            } else {
                for (int g = g0; g < g1; ++g)
                    p.partial[((size_t)ct * PRED_GROUPS + g) * TILE + tid] =
                        sh_grp[g * WM * TILE + tid] + sh_grp[g * WM * TILE + TILE + tid];
            }
            if (split == 0 && j < p.M) {
                double m = 0.0;
#pragma unroll
                for (int o = 0; o < OS; ++o) m += sh_mu[(o * 64 + pj) * 2 + c];
                p.mean[oidx] = m;
            }
        }
        // the next item's shared-memory partials are written only after its first __syncthreads
    }
}

// nsplit > 1: var = k** - sum_g partial[tile][g][col], same order as the in-kernel sum
template <int FAM>
__global__ void __launch_bounds__(TILE) predict_finish_kernel(PredictArgs p) {
    if (p.gate && *p.gate != 0) return;
    const int ct = blockIdx.x, tid = threadIdx.x;
    const int c = tid & 1;
    const int j = FAM == FAM_HELM ? ct * 64 + (tid >> 1) : ct * 128 + tid;
    const long oidx = FAM == FAM_HELM ? (long)c * p.out_stride + j : (long)j;
    if (j >= p.M) return;
    double ss = 0.0;
    for (int g = 0; g < PRED_GROUPS; ++g) ss += p.partial[((size_t)ct * PRED_GROUPS + g) * TILE + tid];
    double v = (FAM == FAM_HELM && c ? p.kss1 : p.kss) - ss;
    v = v < 0.0 ? 0.0 : v;
    p.var[oidx] = v + p.var_add;
}

// + 20 KB: consecutive panels must not sit at the same offset modulo a power of two (every CTA
// streams its panel at the same pace; identical low address bits would pile onto the same L2 slices)
size_t predict_panel_bytes(int npad) { return (size_t)npad * TILE * sizeof(double) + 20480; }

int predict_max_ctas() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

// Split of the row blocks of one column tile over CTAs.  An item costs its share of the DMMA work
// plus a full phase 1 (every CTA of a split tile generates the whole panel), so the time is
// ~ ceil(ntiles nsplit / SMs) * (1 / nsplit + phi) with phi = phase 1 / DMMA time of a whole tile
// (measured with tools/explore.py split at npad = 4096: 1.15 % for the Helmholtz block generator,
// 3.4 % for the scalar family, which needs one exponential per entry; both scale as 1 / npad).  The smallest nsplit in
// {1,2,4,8} within 3 % of the best estimate is taken.
static thread_local int g_force_split = 0;                 // bring-up override, per host thread
void set_predict_split(int s) { g_force_split = (s == 1 || s == 2 || s == 4 || s == 8) ? s : 0; }

int predict_choose_split(long ntiles, int npad, int fam) {
    if (g_force_split) return g_force_split;      // bring-up override (calibration of phi)
    const long sms = predict_max_ctas();
    const double phi = (fam == FAM_HELM ? 47.0 : 140.0) / (double)npad;
    double best = 1e300;
    double r[4];
    for (int i = 0; i < 4; ++i) {
        const long s = 1L << i;
        r[i] = (double)((ntiles * s + sms - 1) / sms) * (1.0 / (double)s + phi);
        if (r[i] < best) best = r[i];
    }
    for (int i = 0; i < 4; ++i)
        if (r[i] <= best * 1.03) return 1 << i;
    return 1;
}

size_t predict_partial_bytes(long ntiles) { return (size_t)ntiles * PRED_GROUPS * TILE * sizeof(double); }

size_t predict_scratch_bytes(int npad, int M, int pts_per_tile) {
    const long ntiles = ((long)M + pts_per_tile - 1) / pts_per_tile;
    const int ns = predict_choose_split(ntiles, npad, pts_per_tile == 64 ? FAM_HELM : FAM_RBF);
    long ctas = predict_max_ctas();
    if (ntiles * ns < ctas) ctas = ntiles * ns;
    return (size_t)ctas * predict_panel_bytes(npad) + (ns > 1 ? predict_partial_bytes(ntiles) : 0);
}

template <int FAM>
static cudaError_t predict_launch(PredictArgs& a, double* scratch, size_t scratch_bytes, cudaStream_t st) {
    static PerDeviceOnce once;
    const int slot = once.pending();
    if (slot >= 0) {
        cudaError_t e = cudaFuncSetAttribute(predict_kernel<FAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        once.done[slot] = true;
    }
    constexpr int PTS = FAM == FAM_HELM ? 64 : 128;
    a.ntiles = (a.M + PTS - 1) / PTS;
    // scratch = [partial sums (nsplit > 1)] [one K* panel per CTA]; a buffer too small for the
    // partials falls back to nsplit = 1, fewer panels only shrink the grid
    a.nsplit = predict_choose_split(a.ntiles, a.npad, FAM);
    size_t pb = predict_partial_bytes(a.ntiles);
    if (a.nsplit > 1 && scratch_bytes < pb + predict_panel_bytes(a.npad)) a.nsplit = 1;
    if (a.nsplit == 1) pb = 0;
    a.partial = scratch;
    a.scratch = scratch + pb / sizeof(double);
    long panels = (long)((scratch_bytes - pb) / predict_panel_bytes(a.npad));
    const long items = (long)a.ntiles * a.nsplit;
    long grid = items;
    if (grid > predict_max_ctas()) grid = predict_max_ctas();
    if (grid > panels) grid = panels;
    if (grid <= 0) return cudaErrorInvalidValue;
    // balance the tail: every CTA gets ceil(items/grid) or one fewer items
    long per = (items + grid - 1) / grid;
    grid = (items + per - 1) / per;
    predict_kernel<FAM><<<(unsigned)grid, WS_THREADS, PRED_SMEM_BYTES, st>>>(a);
    if (a.nsplit > 1) predict_finish_kernel<FAM><<<a.ntiles, TILE, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t predict_fused(const double* Zt, int npad, const double* alpha_int,
                          const double* X, int N, const HelmParams& hp, const double* Xs, int M,
                          long out_stride, double var_add, double* mean, double* var,
                          double* scratch, size_t scratch_bytes, cudaStream_t st, const int* gate) {
    if (M <= 0) return cudaSuccess;
    PredictArgs a{};
    a.gate = gate;
    a.Zt = Zt; a.npad = npad; a.alpha = alpha_int; a.X = X; a.N = N; a.hp = hp;
    a.Xs = Xs; a.M = M; a.out_stride = out_stride;
    a.kss = a.kss1 = hp.tvar * (hp.w_df + hp.w_cf);   // ratio/l_df^2 + (1-ratio)/l_cf^2 (myKernel.py:55-57), times the time variance
    a.var_add = var_add; a.mean = mean; a.var = var;
    return predict_launch<FAM_HELM>(a, scratch, scratch_bytes, st);
}

cudaError_t predict_fused_hsum(const double* Zt, int npad, const double* alpha_int, const double* X, int N,
                               const HsumParams& sp, const double* Xs, int M, long out_stride, double var_add,
                               double* mean, double* var, double* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    PredictArgs a{};
    a.Zt = Zt; a.npad = npad; a.alpha = alpha_int; a.X = X; a.N = N; a.sp = sp; a.use_hsum = 1;
    a.Xs = Xs; a.M = M; a.out_stride = out_stride;
    a.kss = sp.kss0; a.kss1 = sp.kss1;
    a.var_add = var_add; a.mean = mean; a.var = var;
    return predict_launch<FAM_HELM>(a, scratch, scratch_bytes, st);
}

cudaError_t predict_fused_rbf(const double* Zt, int npad, const double* alpha, const double* X, int N,
                              const RbfParams& rp, const double* Xs, int M, double var_add, double* mean,
                              double* var, double* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    PredictArgs a{};
    a.Zt = Zt; a.npad = npad; a.alpha = alpha; a.X = X; a.N = N; a.rp = rp;
    a.Xs = Xs; a.M = M; a.out_stride = M;
    a.kss = a.kss1 = rp.kss;
    a.var_add = var_add; a.mean = mean; a.var = var;
    return predict_launch<FAM_RBF>(a, scratch, scratch_bytes, st);
}

}  // namespace gp2d
