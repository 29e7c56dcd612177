// dense.cuh -- large-N path (N > 32): the propagator is never formed. Each time
// step applies the Taylor polynomial of exp(-i H_n dt) to the whole block of K
// states (forward) or to the GradGenerator block [chi'_1 .. chi'_L, chi] of
// K(L+1) states (backward, reference docs/src/background.md:447-494) -- a chain
// of m dense complex GEMMs  H_n (N x N) * T_{j-1} (N x C)  per step, executed
// with FP64 tensor-core DMMA (mma.sync.m8n8k4.f64) on planar (split re/im) data.
//
// One persistent cooperative kernel per sweep keeps the whole time loop on the
// device: CTA (row tile rt, column part) owns 8 rows x its columns of every
// block; its 8 rows of H_n stay in shared memory for all m terms of a step, its
// slice of the accumulated state stays in shared memory for the whole step, and
// the only global synchronisation is one grid barrier per Taylor term (the
// new term T_j must be complete before any CTA reads it as the next operand).
//
// Layout in HBM (all planar: re plane followed by im plane, row-major, padded
// Np = ceil32(N), Kp = ceil8(K)):
//   Hf [(1+L)][2][Np][Np]   H0, Hc_1..Hc_L          Ha: their adjoints
//   Dm [nD][2][Np][Np]      quadratic-form running cost operator
//   cur  [2][Np][Kp]        current forward block      bcur [2][Np][Cb], Cb = (L+1) Kp
//   T0/T1 [2][Np][Cb]       ping-pong Taylor terms
//   store [(NT+1)][2][Np][Kp]  fw_storage (reference src/workspace.jl:215)
#pragma once
#include "common.cuh"
#include "reduce.cuh"
#include "../../include/grape_b200.h"
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace cgx = cooperative_groups;

constexpr int DENSE_LMAX = 8;
constexpr int DENSE_CGP = 4;      // column groups (of 8 columns) per DMMA pass
constexpr int DENSE_THREADS = 256;

constexpr int KRY_MTMAX = 20;   // compile-time cap of Taylor-term slots per step (Krylov-form backward)

// Krylov-form backward (dense_kry.cuh): per step the Taylor terms of both sweeps are kept in HBM,
//   bh_a = (-i H dt)^a Psi / a!   (forward chain),   ch_b = (+i H^dagger dt)^b chi / b!   (backward chain),
// and the gradient of ALL controls comes from one contraction per step (no mu_l mat-vecs in the chain).
struct KryDev {
    int on;            // term storage allocated and the Krylov-form backward is available for this handle
    int MT;            // term slots per step (<= KRY_MTMAX)
    int TP, TT;        // 64x64 contraction tiles per dimension / per step
    int KC;            // contraction chunk (8 or 16 columns of a term slot)
    int* m_n;          // [NT] Taylor order of every step (kry_plan)
    int* ok;           // [1] 1: every step has s == 0 and m <= MT -> Krylov-form kernels run, block recursion skips
    int* kb;           // [1] number of partial-gradient slabs finalize_grad sums (1 on the Krylov path)
    double* FT;        // [NT][MT][2][Np][Kp]  forward terms, overwritten by e_b = rho sum_a beta(a,b) bh_a
    double* BT;        // [NT][MT][2][Np][Kp]  backward terms, slot 0 = chi(t_n)
    double* kcur;      // [2][Np][Kp]  state of the backward chain
    double* kcur2;     // ping-pong partner (tiled chain)
    double* tilepart;  // [NT][TT][L]
};

struct DenseDev {
    int Np, Kp, Cb, RT, Pf, Pb, MS, CcapF, CcapB, nD;
    const int* kry_ok; // block-recursion backward kernels return at once if *kry_ok (nullptr: always run)
    int mu_smem;   // backward strip kernel keeps its 8 rows of every mu_l^dagger in shared memory
    const double* Hf;
    const double* Ha;
    const double* Dm;
    double hnorm[1 + DENSE_LMAX];
    double* cur;
    double* bcur;
    double* T0;
    double* T1;
    double* store;
    double* jbpart;    // [gridF][Kp]
    const double* tgt; // planar [2][Np][Kp]
    const double* psi0;
};

struct DensePlan {
    DenseDev d;
    KryDev kd;
    size_t kry_smem;
    int gridF, gridB;
    size_t smemF, smemB;
    bool ready;
    bool strip_ok;        // the 8-row strip kernels of this file can run (N, K(L+1) small enough)
    std::string strip_err;
    DensePlan() : kry_smem(0), gridF(0), gridB(0), smemF(0), smemB(0), ready(false), strip_ok(true) { memset(&kd, 0, sizeof kd); }
};

GB_D void dmma884(double (&acc)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(acc[0]), "+d"(acc[1]) : "d"(a), "d"(b));
}

struct DAcc {
    double p1[2], p2[2], q1[2], q2[2];
    GB_D void zero() { p1[0] = p1[1] = p2[0] = p2[1] = q1[0] = q1[1] = q2[0] = q2[1] = 0.0; }
};

// Accumulates, for NG column groups, sum over this warp's k-slice of
//   B[r][k] * X[k][c]   with B = 8 rows (stride bstride, scaled by bscale), X planar global (ld = ldx)
// DMMA mapping: A[m][k] = X[k0+k][col0+m], B[k][n] = Brow[n][k0+k], D[m][n] = out[row n][col m].
// The operand block X was written by other CTAs one grid barrier ago and comes from L2: the loads of UK
// k-steps are issued back to back before the first DMMA so that their latency overlaps.
// NSET independent accumulator sets are used round-robin over the k-steps: a dependent DMMA has ~100 clk of
// latency, and with one 8-column group per warp a single set would be a chain of kslice/4 dependent DMMAs.
template <int NG, bool BSMEM, int UK, int NSET = 1>
GB_D void dense_mma_slice(const double* __restrict__ Bre, const double* __restrict__ Bim, int bstride,
                          double bscale, const double* __restrict__ Xre, const double* __restrict__ Xim,
                          int ldx, const int (&col0)[NG], int ng, int kbeg, int kend, DAcc (&acc)[NG]) {
    static_assert(UK % NSET == 0, "UK must be a multiple of NSET");
    const int lane = threadIdx.x & 31;
    const int lr = lane >> 2, lc = lane & 3;
    DAcc loc[NSET > 1 ? NSET - 1 : 1][NG];
    if (NSET > 1) {
#pragma unroll
        for (int q = 0; q < NSET - 1; ++q)
#pragma unroll
            for (int g = 0; g < NG; ++g) loc[q][g].zero();
    }
    for (int kb = kbeg; kb < kend; kb += 4 * UK) {
        double are[UK][NG], aim[UK][NG];
#pragma unroll
        for (int u = 0; u < UK; ++u) {
            const int k0 = kb + 4 * u;
            if (k0 < kend) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (g < ng) {
                        const size_t off = (size_t)(k0 + lc) * ldx + col0[g] + lr;
                        are[u][g] = __ldcg(&Xre[off]);
                        aim[u][g] = __ldcg(&Xim[off]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UK; ++u) {
            const int k0 = kb + 4 * u;
            if (k0 < kend) {
                double bre, bim;
                if (BSMEM) {
                    bre = bscale * Bre[lr * bstride + k0 + lc];
                    bim = bscale * Bim[lr * bstride + k0 + lc];
                } else {
                    bre = bscale * __ldg(&Bre[(size_t)lr * bstride + k0 + lc]);
                    bim = bscale * __ldg(&Bim[(size_t)lr * bstride + k0 + lc]);
                }
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (g < ng) {
                        DAcc& a = (u % NSET == 0) ? acc[g] : loc[(u % NSET) - (NSET > 1 ? 1 : 0)][g];
                        dmma884(a.p1, are[u][g], bre);
                        dmma884(a.p2, aim[u][g], bim);
                        dmma884(a.q1, are[u][g], bim);
                        dmma884(a.q2, aim[u][g], bre);
                    }
                }
            }
        }
    }
    if (NSET > 1) {
#pragma unroll
        for (int q = 0; q < NSET - 1; ++q)
#pragma unroll
            for (int g = 0; g < NG; ++g)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    acc[g].p1[e] += loc[q][g].p1[e];
                    acc[g].p2[e] += loc[q][g].p2[e];
                    acc[g].q1[e] += loc[q][g].q1[e];
                    acc[g].q2[e] += loc[q][g].q2[e];
                }
    }
}

// Cross-warp (k-split) reduction. After the call, thread t holds the complex result for
// column group g = t/64, row nrow = (t%64)/8, column m = t%8 (valid if g < ng).
template <int NG>
GB_D cplx dense_reduce(double* __restrict__ red, DAcc (&acc)[NG], int ng) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (g < ng) {
            double* rr = red + (size_t)((w * NG + g) * 2) * 64;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                rr[(2 * lc + e) * 8 + lr] = acc[g].p1[e] - acc[g].p2[e];
                rr[64 + (2 * lc + e) * 8 + lr] = acc[g].q1[e] + acc[g].q2[e];
            }
        }
    }
    __syncthreads();
    const int g = threadIdx.x >> 6, idx = threadIdx.x & 63;
    cplx res = mk(0.0, 0.0);
    if (g < ng) {
#pragma unroll
        for (int ww = 0; ww < DENSE_THREADS / 32; ++ww) {
            const double* rr = red + (size_t)((ww * NG + g) * 2) * 64;
            res.x += rr[idx];
            res.y += rr[64 + idx];
        }
    }
    __syncthreads();
    return res;
}

GB_D void dense_plan(const DevP& p, const DenseDev& d, int n, double dt, int& m, int& s) {
    double nrm = d.hnorm[0];
    for (int l = 0; l < p.L; ++l) {
        double a = p.eps[l * p.NT + n];
        if (p.shape) a *= p.shape[l * p.NT + n];
        nrm += fabs(a) * d.hnorm[1 + l];
    }
    vec_plan(nrm * dt, m, s);
}

// rows r0..r0+7 of  H0 + sum_l a_l Hc_l  (or of the adjoints) into shared memory; the 2 x 8 loads of a trip and
// operator are issued before their first use (this loop is pure load latency otherwise)
GB_D void dense_form_H(const DevP& p, const double* __restrict__ Hall, int Np, int MS, int r0, int n,
                       double* __restrict__ Hs_re, double* __restrict__ Hs_im) {
    const size_t plane = (size_t)Np * Np;
    constexpr int U = 8;
    const int tot = 8 * Np;
    const double* __restrict__ Hrow = Hall + (size_t)r0 * Np;   // the 8 rows are contiguous: offset = e
    for (int e0 = threadIdx.x; e0 < tot; e0 += U * DENSE_THREADS) {
        double hr[U], hi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * DENSE_THREADS;
            hr[u] = e < tot ? __ldg(&Hrow[e]) : 0.0;
            hi[u] = e < tot ? __ldg(&Hrow[plane + e]) : 0.0;
        }
        for (int l = 0; l < p.L; ++l) {
            double al = p.eps[l * p.NT + n];
            if (p.shape) al *= p.shape[l * p.NT + n];
            const double* __restrict__ Hl = Hrow + (size_t)(1 + l) * 2 * plane;
            double tr[U], ti[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * DENSE_THREADS;
                tr[u] = e < tot ? __ldg(&Hl[e]) : 0.0;
                ti[u] = e < tot ? __ldg(&Hl[plane + e]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                hr[u] = fma(al, tr[u], hr[u]);
                hi[u] = fma(al, ti[u], hi[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * DENSE_THREADS;
            if (e < tot) {
                const int r = e / Np, k = e - r * Np;
                Hs_re[r * MS + k] = hr[u];
                Hs_im[r * MS + k] = hi[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Chain kernel: forward sweep (BWD = false; reference src/optimize.jl:720-751) and the chi chain of the
// Krylov-form backward (BWD = true; chi <- exp(+i H^dagger dt) chi going down in n, plus the running-cost
// inhomogeneity of optimize.jl:897-908). Same Taylor recursion on a block of K states in both directions.
// When the step qualifies for the Krylov form (s == 0, m <= MT) every Taylor term goes to its own HBM slot
// (FT / BT) instead of the T0/T1 ping-pong, at no extra traffic.
// ---------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(DENSE_THREADS, 1) dense_chain(DevP p, DenseDev d, KryDev kd) {
    if (BWD && !(*kd.ok)) return;   // uniform over the grid
    extern __shared__ __align__(16) double dsm[];
    const int Np = d.Np, Kp = d.Kp, MS = d.MS, NT = p.NT;
    const int rt = blockIdx.x / d.Pf, part = blockIdx.x % d.Pf;
    // One cooperative_groups grid barrier per Taylor term: ncu attributes ~40 % of this kernel to it (arrival skew
    // after every CTA pulls its whole operand block from L2 at the same instant + the barrier round trips).  Two
    // replacements were measured on C4 and were SLOWER (profiles/r1_s5_*, r1_s7_*): a two-level counter barrier
    // with one domain per column part (+0.6 us per term) and per-producer release/acquire counters polled by every
    // consumer warp (+1.5 us per term: thousands of pollers on two L2 lines).
    cgx::grid_group grid = cgx::this_grid();
    const int CGtot = Kp / 8;
    const int cg0 = (part * CGtot) / d.Pf, cg1 = ((part + 1) * CGtot) / d.Pf;
    const int ncols = (cg1 - cg0) * 8, cbeg = cg0 * 8, Ccap = d.CcapF;
    const int r0 = rt * 8;
    double* Hs_re = dsm;
    double* Hs_im = Hs_re + 8 * MS;
    double* acc_re = Hs_im + 8 * MS;
    double* acc_im = acc_re + 8 * Ccap;
    double* red = acc_im + 8 * Ccap;
    double* jb_s = red + (DENSE_THREADS / 32) * DENSE_CGP * 128;
    const size_t splane = (size_t)Np * Kp;
    const int w = threadIdx.x >> 5;
    const int kslice = Np / 8, kbeg = w * kslice, kend = kbeg + kslice;
    const bool gb = BWD ? (p.gb_kind != 0 && p.lambda_b != 0.0) : (p.gb_kind != 0);
    const size_t hplane = (size_t)Np * Np;
    double* state = BWD ? kd.kcur : d.cur;
    const double* Hall = BWD ? d.Ha : d.Hf;
    double* terms = BWD ? kd.BT : kd.FT;

    for (int e = threadIdx.x; e < 8 * ncols; e += DENSE_THREADS) {
        const int r = e / ncols, c = e % ncols;
        const size_t off = (size_t)(r0 + r) * Kp + cbeg + c;
        const double vr = state[off], vi = state[splane + off];
        acc_re[r * Ccap + c] = vr;
        acc_im[r * Ccap + c] = vi;
        if (BWD) {   // slot 0 of the last step = chi(T)
            double* s0 = terms + (size_t)(NT - 1) * kd.MT * 2 * splane;
            s0[off] = vr;
            s0[splane + off] = vi;
        }
    }
    for (int c = threadIdx.x; c < Ccap; c += DENSE_THREADS) jb_s[c] = 0.0;
    __syncthreads();

    // J_b contribution of the state currently in `cur` (all rows) / acc (my rows), weight wgt
    auto gb_point = [&](double wgt) {
        for (int pg = cg0; pg < cg1; pg += DENSE_CGP) {
            const int ng = min(DENSE_CGP, cg1 - pg);
            int col0[DENSE_CGP];
            DAcc acc[DENSE_CGP];
#pragma unroll
            for (int g = 0; g < DENSE_CGP; ++g) { col0[g] = (pg + g) * 8; acc[g].zero(); }
            // D is shared (nD == 1) on the dense path
            dense_mma_slice<DENSE_CGP, false, 2>(d.Dm + (size_t)r0 * Np, d.Dm + hplane + (size_t)r0 * Np, Np, 1.0,
                                                 d.cur, d.cur + splane, Kp, col0, ng, kbeg, kend, acc);
            const cplx res = dense_reduce<DENSE_CGP>(red, acc, ng);
            const int g = threadIdx.x >> 6, idx = threadIdx.x & 63, nrow = idx >> 3, mc = idx & 7;
            // reuse `red` as scratch for the row reduction
            double v = 0.0;
            if (g < ng) {
                const int c = (pg + g) * 8 + mc - cbeg;
                v = acc_re[nrow * Ccap + c] * res.x + acc_im[nrow * Ccap + c] * res.y;
            }
            red[threadIdx.x] = v;
            __syncthreads();
            if (g < ng && nrow == 0) {
                double sacc = 0.0;
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) sacc += red[g * 64 + rr * 8 + mc];
                jb_s[(pg + g) * 8 + mc - cbeg] += wgt * sacc;
            }
            __syncthreads();
        }
    };

    for (int it = 0; it < NT; ++it) {
        const int n = BWD ? NT - 1 - it : it;
        const double dt = p.tlist[n + 1] - p.tlist[n];
        if (!BWD && gb) {
            const double wgt = n == 0 ? 0.5 * (p.tlist[1] - p.tlist[0]) : 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]);
            gb_point(wgt);
        }
        dense_form_H(p, Hall, Np, MS, r0, n, Hs_re, Hs_im);
        int m, s;
        dense_plan(p, d, n, dt, m, s);
        if (!BWD && p.grad_method != 0 && p.taylor_check && m > p.taylor_max_order && blockIdx.x == 0 && threadIdx.x == 0)
            p.flags->taylor_fail = 1;
        const bool kry = kd.on && s == 0 && m <= kd.MT;
        double* slots = kry ? terms + (size_t)n * kd.MT * 2 * splane : nullptr;
        __syncthreads();
        const int nsub = 1 << s;
        const double dts = dt / nsub;
        for (int sub = 0; sub < nsub; ++sub) {
            for (int j = 1; j <= m; ++j) {
                const double* src = j == 1 ? state : (kry ? slots + (size_t)(j - 1) * 2 * splane : ((j - 1) & 1 ? d.T1 : d.T0));
                double* dst = kry ? slots + (size_t)j * 2 * splane : ((j & 1) ? d.T1 : d.T0);
                const double x = dts / j;
                        for (int pg = cg0; pg < cg1; pg += DENSE_CGP) {
                    const int ng = min(DENSE_CGP, cg1 - pg);
                    cplx res;
                    if (ng == 1) {   // 8 columns per CTA (few trajectories): deeper load batches
                        int col0[1] = {pg * 8};
                        DAcc acc[1];
                        acc[0].zero();
                        dense_mma_slice<1, true, 8, 4>(Hs_re, Hs_im, MS, 1.0, src, src + splane, Kp, col0, 1, kbeg, kend, acc);
                        res = dense_reduce<1>(red, acc, 1);
                    } else {
                        int col0[DENSE_CGP];
                        DAcc acc[DENSE_CGP];
#pragma unroll
                        for (int g = 0; g < DENSE_CGP; ++g) { col0[g] = (pg + g) * 8; acc[g].zero(); }
                        dense_mma_slice<DENSE_CGP, true, 2>(Hs_re, Hs_im, MS, 1.0, src, src + splane, Kp, col0, ng, kbeg, kend, acc);
                        res = dense_reduce<DENSE_CGP>(red, acc, ng);
                    }
                    const int g = threadIdx.x >> 6, idx = threadIdx.x & 63, nrow = idx >> 3, mc = idx & 7;
                    if (g < ng) {
                        // forward: t = (-i x) * res ; backward: t = (+i x) * res
                        const double tr = BWD ? -x * res.y : x * res.y;
                        const double ti = BWD ? x * res.x : -x * res.x;
                        const int cglob = (pg + g) * 8 + mc;
                        if (j < m) {
                            dst[(size_t)(r0 + nrow) * Kp + cglob] = tr;
                            dst[splane + (size_t)(r0 + nrow) * Kp + cglob] = ti;
                        }
                        acc_re[nrow * Ccap + cglob - cbeg] += tr;
                        acc_im[nrow * Ccap + cglob - cbeg] += ti;
                    }
                }
                if (j < m) grid.sync();
            }
            __syncthreads();
            const bool last = sub == nsub - 1;
            if (BWD && last && gb && n > 0) {
                // chi += lambda_b * 0.5 (t_{n+1} - t_{n-1}) / rho * xi(Psi(t_{n-1})), xi = -D Psi  (optimize.jl:897-908)
                const double* st = d.store + (size_t)n * 2 * splane;
                const double f = p.lambda_b * 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]);
                for (int pg = cg0; pg < cg1; pg += DENSE_CGP) {
                    const int ng = min(DENSE_CGP, cg1 - pg);
                    int col0[DENSE_CGP];
                    DAcc acc[DENSE_CGP];
#pragma unroll
                    for (int g = 0; g < DENSE_CGP; ++g) { col0[g] = (pg + g) * 8; acc[g].zero(); }
                    dense_mma_slice<DENSE_CGP, false, 2>(d.Dm + (size_t)r0 * Np, d.Dm + hplane + (size_t)r0 * Np, Np, 1.0,
                                                         st, st + splane, Kp, col0, ng, kbeg, kend, acc);
                    const cplx res = dense_reduce<DENSE_CGP>(red, acc, ng);
                    const int g = threadIdx.x >> 6, idx = threadIdx.x & 63, nrow = idx >> 3, mc = idx & 7;
                    if (g < ng) {
                        const int k = (pg + g) * 8 + mc;
                        if (k < p.K) {
                            const double fk = f / p.rho[k];
                            acc_re[nrow * Ccap + k - cbeg] -= fk * res.x;
                            acc_im[nrow * Ccap + k - cbeg] -= fk * res.y;
                        }
                    }
                }
                __syncthreads();
            }
            for (int e = threadIdx.x; e < 8 * ncols; e += DENSE_THREADS) {
                const int r = e / ncols, c = e % ncols;
                const size_t off = (size_t)(r0 + r) * Kp + cbeg + c;
                const double vr = acc_re[r * Ccap + c], vi = acc_im[r * Ccap + c];
                state[off] = vr;
                state[splane + off] = vi;
                if (last) {
                    if (!BWD) {
                        double* st = d.store + (size_t)(n + 1) * 2 * splane;
                        __stcs(&st[off], vr);
                        __stcs(&st[splane + off], vi);
                    } else if (n > 0) {   // slot 0 of the next (earlier) step = chi(t_{n-1})
                        double* s0 = terms + (size_t)(n - 1) * kd.MT * 2 * splane;
                        s0[off] = vr;
                        s0[splane + off] = vi;
                    }
                }
            }
            grid.sync();
        }
    }
    if (!BWD && gb) {
        gb_point(0.5 * (p.tlist[NT] - p.tlist[NT - 1]));
        for (int c = threadIdx.x; c < ncols; c += DENSE_THREADS)
            d.jbpart[(size_t)blockIdx.x * Kp + cbeg + c] = jb_s[c];
    }
}

// ---------------------------------------------------------------------------
// Backward sweep fused with the gradient contraction
// (reference src/optimize.jl:880-911; GradGenerator block, docs/src/background.md:467-477)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(DENSE_THREADS, 1) dense_backward(DevP p, DenseDev d) {
    if (d.kry_ok && *d.kry_ok) return;   // the Krylov-form kernels (dense_kry.cuh) serve this call; uniform over the grid
    cgx::grid_group grid = cgx::this_grid();
    extern __shared__ __align__(16) double dsm[];
    const int Np = d.Np, Kp = d.Kp, Cb = d.Cb, MS = d.MS, NT = p.NT, L = p.L;
    const int rt = blockIdx.x / d.Pb, part = blockIdx.x % d.Pb;
    const int CGtot = Cb / 8;
    const int cg0 = (part * CGtot) / d.Pb, cg1 = ((part + 1) * CGtot) / d.Pb;
    const int ncols = (cg1 - cg0) * 8, cbeg = cg0 * 8, Ccap = d.CcapB;
    const int r0 = rt * 8;
    double* Hs_re = dsm;
    double* Hs_im = Hs_re + 8 * MS;
    double* acc_re = Hs_im + 8 * MS;
    double* acc_im = acc_re + 8 * Ccap;
    double* red = acc_im + 8 * Ccap;
    double* s_buf = red + (DENSE_THREADS / 32) * DENSE_CGP * 128;   // 32 doubles for block_sum
    double* Mu_s = s_buf + 64;                                      // [L][2][8][MS] if d.mu_smem
    const size_t bplane = (size_t)Np * Cb, splane = (size_t)Np * Kp, hplane = (size_t)Np * Np;
    const int w = threadIdx.x >> 5;
    const int kslice = Np / 8, kbeg = w * kslice, kend = kbeg + kslice;
    const bool gb = p.gb_kind != 0 && p.lambda_b != 0.0;
    const int chi0 = L * Kp;   // first column of the chi block

    for (int e = threadIdx.x; e < 8 * ncols; e += DENSE_THREADS) {
        const int r = e / ncols, c = e % ncols;
        acc_re[r * Ccap + c] = d.bcur[(size_t)(r0 + r) * Cb + cbeg + c];
        acc_im[r * Ccap + c] = d.bcur[bplane + (size_t)(r0 + r) * Cb + cbeg + c];
    }
    if (d.mu_smem) {   // the control operators do not depend on the time step: staged once
        for (int e = threadIdx.x; e < L * 2 * 8 * Np; e += DENSE_THREADS) {
            const int k = e % Np, r = (e / Np) % 8, pl = (e / (8 * Np)) % 2, l = e / (16 * Np);
            Mu_s[((size_t)(l * 2 + pl) * 8 + r) * MS + k] =
                d.Ha[(size_t)(1 + l) * 2 * hplane + (size_t)pl * hplane + (size_t)(r0 + r) * Np + k];
        }
    }
    __syncthreads();

    for (int n = NT - 1; n >= 0; --n) {
        const double dt = p.tlist[n + 1] - p.tlist[n];
        dense_form_H(p, d.Ha, Np, MS, r0, n, Hs_re, Hs_im);
        int m, s;
        dense_plan(p, d, n, dt, m, s);
        __syncthreads();
        const int nsub = 1 << s;
        const double dts = dt / nsub;
        for (int sub = 0; sub < nsub; ++sub) {
            for (int j = 1; j <= m; ++j) {
                const double* src = j == 1 ? d.bcur : ((j - 1) & 1 ? d.T1 : d.T0);
                double* dst = (j & 1) ? d.T1 : d.T0;
                const double x = dts / j;
                for (int pg = cg0; pg < cg1; pg += DENSE_CGP) {
                    const int ng = min(DENSE_CGP, cg1 - pg);
                    int col0[DENSE_CGP];
                    DAcc acc[DENSE_CGP];
#pragma unroll
                    for (int g = 0; g < DENSE_CGP; ++g) { col0[g] = (pg + g) * 8; acc[g].zero(); }
                    // H^dagger * [chi'_l .. chi]
                    dense_mma_slice<DENSE_CGP, true, 4>(Hs_re, Hs_im, MS, 1.0, src, src + bplane, Cb, col0, ng, kbeg, kend, acc);
                    // + mu_l^dagger * chi  into the chi'_l columns
#pragma unroll
                    for (int g = 0; g < DENSE_CGP; ++g) {
                        if (g < ng) {
                            const int l = ((pg + g) * 8) / Kp;
                            if (l < L) {
                                int c1[1] = {chi0 + (pg + g) * 8 - l * Kp};
                                DAcc a1[1];
                                a1[0] = acc[g];
                                const double sl = p.shape ? p.shape[l * NT + n] : 1.0;
                                if (d.mu_smem) {
                                    const double* Ml = Mu_s + (size_t)(l * 2) * 8 * MS;
                                    dense_mma_slice<1, true, 8>(Ml, Ml + 8 * MS, MS, sl, src, src + bplane, Cb, c1, 1, kbeg, kend, a1);
                                } else {
                                    const double* Bl = d.Ha + (size_t)(1 + l) * 2 * hplane + (size_t)r0 * Np;
                                    dense_mma_slice<1, false, 8>(Bl, Bl + hplane, Np, sl, src, src + bplane, Cb, c1, 1, kbeg, kend, a1);
                                }
                                acc[g] = a1[0];
                            }
                        }
                    }
                    const cplx res = dense_reduce<DENSE_CGP>(red, acc, ng);
                    const int g = threadIdx.x >> 6, idx = threadIdx.x & 63, nrow = idx >> 3, mc = idx & 7;
                    if (g < ng) {
                        // t = (+i x) * res
                        const double tr = -x * res.y, ti = x * res.x;
                        const int cglob = (pg + g) * 8 + mc;
                        if (j < m) {
                            dst[(size_t)(r0 + nrow) * Cb + cglob] = tr;
                            dst[bplane + (size_t)(r0 + nrow) * Cb + cglob] = ti;
                        }
                        acc_re[nrow * Ccap + cglob - cbeg] += tr;
                        acc_im[nrow * Ccap + cglob - cbeg] += ti;
                    }
                }
                if (j < m) grid.sync();
            }
            __syncthreads();
            const bool last = sub == nsub - 1;
            if (last) {
                const double* st = d.store + (size_t)n * 2 * splane;    // Psi(t_{n-1}) 1-based = storage index n
                // gradient partials: tau_grad[k][n,l] = rho_k <chi'_lk | Psi_k(t_{n-1})>  (optimize.jl:893-895)
                for (int l = 0; l < L; ++l) {
                    double v[1] = {0.0};
                    for (int e = threadIdx.x; e < 8 * ncols; e += DENSE_THREADS) {
                        const int r = e / ncols, c = e % ncols;
                        const int cglob = cbeg + c;
                        if (cglob >= l * Kp && cglob < (l + 1) * Kp) {
                            const int k = cglob - l * Kp;
                            if (k < p.K) {
                                const size_t so = (size_t)(r0 + r) * Kp + k;
                                v[0] += p.rho[k] * (acc_re[r * Ccap + c] * __ldg(&st[so]) + acc_im[r * Ccap + c] * __ldg(&st[splane + so]));
                            }
                        }
                    }
                    block_sum<1>(v, s_buf);
                    if (threadIdx.x == 0) p.partial[(size_t)blockIdx.x * L * NT + (size_t)l * NT + n] = v[0];
                }
                // resetgradvec!  (optimize.jl:896)
                for (int e = threadIdx.x; e < 8 * ncols; e += DENSE_THREADS) {
                    const int r = e / ncols, c = e % ncols;
                    if (cbeg + c < chi0) { acc_re[r * Ccap + c] = 0.0; acc_im[r * Ccap + c] = 0.0; }
                }
                __syncthreads();
                // chi += lambda_b * 0.5 (t_{n+1} - t_{n-1}) / rho * xi(Psi(t_{n-1})), xi = -D Psi  (optimize.jl:897-908)
                if (gb && n > 0) {
                    const double f = p.lambda_b * 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]);
                    const int pg0 = max(cg0, chi0 / 8);
                    for (int pg = pg0; pg < cg1; pg += DENSE_CGP) {
                        const int ng = min(DENSE_CGP, cg1 - pg);
                        int col0[DENSE_CGP];
                        DAcc acc[DENSE_CGP];
#pragma unroll
                        for (int g = 0; g < DENSE_CGP; ++g) { col0[g] = (pg + g) * 8 - chi0; acc[g].zero(); }
                        dense_mma_slice<DENSE_CGP, false, 4>(d.Dm + (size_t)r0 * Np, d.Dm + hplane + (size_t)r0 * Np, Np, 1.0,
                                                          st, st + splane, Kp, col0, ng, kbeg, kend, acc);
                        const cplx res = dense_reduce<DENSE_CGP>(red, acc, ng);
                        const int g = threadIdx.x >> 6, idx = threadIdx.x & 63, nrow = idx >> 3, mc = idx & 7;
                        if (g < ng) {
                            const int cglob = (pg + g) * 8 + mc;
                            const int k = cglob - chi0;
                            if (k < p.K) {
                                const double fk = f / p.rho[k];
                                acc_re[nrow * Ccap + cglob - cbeg] -= fk * res.x;
                                acc_im[nrow * Ccap + cglob - cbeg] -= fk * res.y;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
            for (int e = threadIdx.x; e < 8 * ncols; e += DENSE_THREADS) {
                const int r = e / ncols, c = e % ncols;
                const size_t off = (size_t)(r0 + r) * Cb + cbeg + c;
                d.bcur[off] = acc_re[r * Ccap + c];
                d.bcur[bplane + off] = acc_im[r * Ccap + c];
            }
            grid.sync();
        }
    }
}

// tau_k = <tgt_k | Psi_k(T)>   (optimize.jl:752-753); also sums the J_b partials over CTAs
__global__ void __launch_bounds__(256) dense_tau(DevP p, DenseDev d, int gridF) {
    __shared__ double s_buf[64];
    const int k = blockIdx.x;
    const size_t splane = (size_t)d.Np * d.Kp;
    double v[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
        const size_t off = (size_t)i * d.Kp + k;
        const double tr = d.tgt[off], ti = d.tgt[splane + off];
        const double xr = d.cur[off], xi = d.cur[splane + off];
        v[0] += tr * xr + ti * xi;
        v[1] += tr * xi - ti * xr;
    }
    block_sum<2>(v, s_buf);
    if (threadIdx.x == 0) {
        p.tau[k] = mk(v[0], v[1]);
        double jb = 0.0;
        if (p.gb_kind)
            for (int b = 0; b < gridF; ++b) jb += d.jbpart[(size_t)b * d.Kp + k];
        p.jb[k] = jb;
    }
}

// chi_k(T) boundary condition, normalisation, initial backward block  (optimize.jl:845-869, 878)
__global__ void __launch_bounds__(256) dense_boundary(DevP p, DenseDev d, const cplx* __restrict__ chi_host, double* __restrict__ kcur) {
    __shared__ double s_buf[32];
    __shared__ double s_rho;
    const int k = blockIdx.x;
    const int Np = d.Np, Kp = d.Kp, Cb = d.Cb, N = p.N;
    const size_t splane = (size_t)Np * Kp, bplane = (size_t)Np * Cb, hplane = (size_t)Np * Np;
    const bool gb = p.gb_kind != 0 && p.lambda_b != 0.0;
    cplx c = mk(0.0, 0.0);
    if (!chi_host) {
        const double w = p.w ? p.w[k] : 1.0;
        const double Kg = (double)p.Kglobal;
        if (p.functional == 0) c = mk(w * p.sums[0] / (Kg * Kg), w * p.sums[1] / (Kg * Kg));
        else if (p.functional == 1) c = mk(w / (2.0 * Kg), 0.0);
        else { cplx t = p.tau[k]; c = mk(w * t.x / Kg, w * t.y / Kg); }
    }
    const double f = gb ? p.lambda_b * (p.tlist[p.NT] - p.tlist[p.NT - 1]) * 0.5 : 0.0;
    double v[1] = {0.0};
    // pass 1: un-normalised chi rows into bcur chi block
    for (int i = threadIdx.x; i < Np; i += blockDim.x) {
        cplx x = mk(0.0, 0.0);
        if (i < N) {
            if (chi_host) x = chi_host[(size_t)k * N + i];
            else x = cmul(c, mk(d.tgt[(size_t)i * Kp + k], d.tgt[splane + (size_t)i * Kp + k]));
            if (gb) {
                double tr = 0.0, ti = 0.0;
                for (int j = 0; j < N; ++j) {
                    const double dr = d.Dm[(size_t)i * Np + j], di = d.Dm[hplane + (size_t)i * Np + j];
                    const double xr = d.cur[(size_t)j * Kp + k], xi = d.cur[splane + (size_t)j * Kp + k];
                    tr += dr * xr - di * xi;
                    ti += dr * xi + di * xr;
                }
                x.x -= f * tr;
                x.y -= f * ti;
            }
        }
        d.bcur[(size_t)i * Cb + p.L * Kp + k] = x.x;
        d.bcur[bplane + (size_t)i * Cb + p.L * Kp + k] = x.y;
        v[0] += cnorm2(x);
    }
    block_sum<1>(v, s_buf);
    if (threadIdx.x == 0) {
        double rho = sqrt(v[0]);
        if (!(rho >= p.chi_min_norm)) {
            if (atomicCAS(&p.flags->chi_bad_k, 0, k + 1) == 0) p.flags->chi_bad_rho = rho;
            rho = 1.0;
        }
        p.rho[k] = rho;
        s_rho = rho;
    }
    __syncthreads();
    const double ir = 1.0 / s_rho;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) {
        const size_t off = (size_t)i * Cb + p.L * Kp + k;
        const double xr = d.bcur[off] * ir, xi = d.bcur[bplane + off] * ir;
        d.bcur[off] = xr;
        d.bcur[bplane + off] = xi;
        if (kcur) {   // K-column state of the Krylov-form backward chain
            kcur[(size_t)i * Kp + k] = xr;
            kcur[splane + (size_t)i * Kp + k] = xi;
        }
        if (i < N) p.chiT[(size_t)k * N + i] = mk(xr, xi);
    }
}

__global__ void dense_gather_states_k(const double* __restrict__ store, cplx* __restrict__ out, int Np, int Kp, int N, int NT, int k) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (NT + 1) * N) {
        const int n = idx / N, i = idx % N;
        const size_t splane = (size_t)Np * Kp;
        const double* st = store + (size_t)n * 2 * splane;
        out[idx] = mk(st[(size_t)i * Kp + k], st[splane + (size_t)i * Kp + k]);
    }
}
__global__ void dense_gather_final_k(const double* __restrict__ cur, cplx* __restrict__ out, int Np, int Kp, int N, int K) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < K * N) {
        const int k = idx / N, i = idx % N;
        const size_t splane = (size_t)Np * Kp;
        out[idx] = mk(cur[(size_t)i * Kp + k], cur[splane + (size_t)i * Kp + k]);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// spectral norm (largest singular value) by power iteration on M^dagger M; planar row-major input
inline double dense_norm2(const double* re, const double* im, int N, int ld) {
    std::vector<double> vr(N), vi(N, 0.0), wr(N), wi(N);
    for (int i = 0; i < N; ++i) vr[i] = 1.0 + 0.37 * std::sin(1.7 * i + 0.3);
    double sigma = 0.0;
    for (int it = 0; it < 60; ++it) {
        double nv = 0.0;
        for (int i = 0; i < N; ++i) nv += vr[i] * vr[i] + vi[i] * vi[i];
        nv = std::sqrt(nv);
        if (nv == 0.0) return 0.0;
        for (int i = 0; i < N; ++i) { vr[i] /= nv; vi[i] /= nv; }
        for (int i = 0; i < N; ++i) {   // w = M v
            double sr = 0.0, si = 0.0;
            const double* rr = re + (size_t)i * ld;
            const double* ri = im + (size_t)i * ld;
            for (int j = 0; j < N; ++j) { sr += rr[j] * vr[j] - ri[j] * vi[j]; si += rr[j] * vi[j] + ri[j] * vr[j]; }
            wr[i] = sr; wi[i] = si;
        }
        double nw = 0.0;
        for (int i = 0; i < N; ++i) nw += wr[i] * wr[i] + wi[i] * wi[i];
        sigma = std::sqrt(nw);
        std::fill(vr.begin(), vr.end(), 0.0);
        std::fill(vi.begin(), vi.end(), 0.0);
        for (int i = 0; i < N; ++i) {   // v = M^dagger w
            const double* rr = re + (size_t)i * ld;
            const double* ri = im + (size_t)i * ld;
            const double ar = wr[i], ai = wi[i];
            for (int j = 0; j < N; ++j) { vr[j] += rr[j] * ar + ri[j] * ai; vi[j] += rr[j] * ai - ri[j] * ar; }
        }
    }
    return sigma;
}

inline void dense_destroy(DensePlan&) {}

inline int dense_setup(DensePlan& dp, DevP& p, const grape_b200_problem* desc, std::vector<void*>& allocs, std::string& err) {
    const int K = p.K, N = p.N, L = p.L, NT = p.NT;
    if (p.G != 1) { err = "dense path (N > 32) supports one shared generator (G == 1)"; return GRAPE_B200_EINVAL; }
    if (L > DENSE_LMAX) { err = "dense path supports at most 8 controls"; return GRAPE_B200_EINVAL; }
    if (p.gb_kind && p.gb_nD != 1) { err = "dense path supports one shared g_b operator D"; return GRAPE_B200_EINVAL; }
    DenseDev& d = dp.d;
    memset(&d, 0, sizeof d);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int Np = (N + 31) / 32 * 32;
    int Kp = (K + 7) / 8 * 8;
    {
        // can the 8-row strip kernels serve this problem (one row strip per SM, tile in shared memory)?  If not, the
        // tiled kernels (dense2.cuh) must: they need the trajectory block padded to their 16-column tile.
        const int RT = Np / 8, MS = Np + 4, Cb8 = (L + 1) * Kp;
        const int Pf = std::max(1, std::min(std::max(1, sms / RT), Kp / 8)), Pb = std::max(1, std::min(std::max(1, sms / RT), Cb8 / 8));
        const size_t CcF = ((Kp / 8 + Pf - 1) / Pf) * 8, CcB = ((Cb8 / 8 + Pb - 1) / Pb) * 8;
        const size_t redB = (size_t)(DENSE_THREADS / 32) * DENSE_CGP * 128;
        const size_t sF = sizeof(double) * (16 * (size_t)MS + 16 * CcF + redB + CcF + 8);
        const size_t sB = sizeof(double) * (16 * (size_t)MS + 16 * CcB + redB + 64);
        if (RT > sms || sF > 227 * 1024 || sB > 227 * 1024) Kp = (K + 15) / 16 * 16;
    }
    const int Cb = (L + 1) * Kp;
    d.Np = Np; d.Kp = Kp; d.Cb = Cb; d.RT = Np / 8; d.MS = Np + 4; d.nD = p.gb_nD;
    if (d.RT > sms) { dp.strip_ok = false; dp.strip_err = "dense strip kernels need ceil32(N)/8 <= number of SMs (N <= 1184 on B200)"; }
    d.Pf = std::max(1, std::min(std::max(1, sms / d.RT), Kp / 8));
    d.Pb = std::max(1, std::min(std::max(1, sms / d.RT), Cb / 8));
    d.CcapF = ((Kp / 8 + d.Pf - 1) / d.Pf) * 8;
    d.CcapB = ((Cb / 8 + d.Pb - 1) / d.Pb) * 8;
    dp.gridF = d.RT * d.Pf;
    dp.gridB = d.RT * d.Pb;
    const size_t hplane = (size_t)Np * Np;
    auto upd = [&](const std::vector<double>& b, const double** dst) -> int {
        void* q = nullptr;
        if (cudaMalloc(&q, b.size() * sizeof(double)) != cudaSuccess) { err = "cudaMalloc failed (dense path)"; return GRAPE_B200_ECUDA; }
        allocs.push_back(q);
        if (cudaMemcpy(q, b.data(), b.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) { err = "cudaMemcpy failed"; return GRAPE_B200_ECUDA; }
        *dst = static_cast<const double*>(q);
        return 0;
    };
    auto ald = [&](double** dst, size_t n) -> int {
        void* q = nullptr;
        if (cudaMalloc(&q, n * sizeof(double)) != cudaSuccess) { err = "cudaMalloc failed (dense path): out of device memory?"; return GRAPE_B200_ECUDA; }
        allocs.push_back(q);
        if (cudaMemset(q, 0, n * sizeof(double)) != cudaSuccess) { err = "cudaMemset failed"; return GRAPE_B200_ECUDA; }
        *dst = static_cast<double*>(q);
        return 0;
    };
    // operators: column-major ABI (element (i,j) at j*N+i) -> planar row-major padded, plus adjoints
    std::vector<double> hf((size_t)(1 + L) * 2 * hplane, 0.0), ha((size_t)(1 + L) * 2 * hplane, 0.0);
    for (int q = 0; q <= L; ++q) {
        const double* src = q == 0 ? desc->H0 : desc->Hc + 2 * (size_t)(q - 1) * N * N;
        double* fr = hf.data() + (size_t)q * 2 * hplane;
        double* ar = ha.data() + (size_t)q * 2 * hplane;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                const double re = src[2 * ((size_t)j * N + i)], im = src[2 * ((size_t)j * N + i) + 1];
                fr[(size_t)i * Np + j] = re;
                fr[hplane + (size_t)i * Np + j] = im;
                ar[(size_t)j * Np + i] = re;            // adjoint: conj transpose
                ar[hplane + (size_t)j * Np + i] = -im;
            }
        d.hnorm[q] = 1.05 * dense_norm2(fr, fr + hplane, N, Np);
    }
    int rc;
    if ((rc = upd(hf, &d.Hf))) return rc;
    if ((rc = upd(ha, &d.Ha))) return rc;
    if (p.gb_kind) {
        std::vector<double> dm(2 * hplane, 0.0);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                dm[(size_t)i * Np + j] = desc->gb_D[2 * ((size_t)j * N + i)];
                dm[hplane + (size_t)i * Np + j] = desc->gb_D[2 * ((size_t)j * N + i) + 1];
            }
        if ((rc = upd(dm, &d.Dm))) return rc;
    }
    const size_t splane = (size_t)Np * Kp, bplane = (size_t)Np * Cb;
    {
        std::vector<double> b(2 * splane, 0.0);
        for (int k = 0; k < K; ++k)
            for (int i = 0; i < N; ++i) {
                b[(size_t)i * Kp + k] = desc->psi0[2 * ((size_t)k * N + i)];
                b[splane + (size_t)i * Kp + k] = desc->psi0[2 * ((size_t)k * N + i) + 1];
            }
        if ((rc = upd(b, &d.psi0))) return rc;
        std::fill(b.begin(), b.end(), 0.0);
        for (int k = 0; k < K; ++k)
            for (int i = 0; i < N; ++i) {
                b[(size_t)i * Kp + k] = desc->tgt[2 * ((size_t)k * N + i)];
                b[splane + (size_t)i * Kp + k] = desc->tgt[2 * ((size_t)k * N + i) + 1];
            }
        if ((rc = upd(b, &d.tgt))) return rc;
    }
    if ((rc = ald(&d.cur, 2 * splane))) return rc;
    if ((rc = ald(&d.bcur, 2 * bplane))) return rc;
    if ((rc = ald(&d.T0, 2 * bplane))) return rc;
    if ((rc = ald(&d.T1, 2 * bplane))) return rc;
    if ((rc = ald(&d.store, (size_t)(NT + 1) * 2 * splane))) return rc;
    if ((rc = ald(&d.jbpart, (size_t)dp.gridF * Kp))) return rc;
    p.KB = dp.gridB;
    if ((rc = ald(&p.partial, (size_t)dp.gridB * L * NT))) return rc;
    const size_t redB = (size_t)(DENSE_THREADS / 32) * DENSE_CGP * 128;
    dp.smemF = sizeof(double) * (16 * (size_t)d.MS + 16 * (size_t)d.CcapF + redB + d.CcapF + 8);
    dp.smemB = sizeof(double) * (16 * (size_t)d.MS + 16 * (size_t)d.CcapB + redB + 64);
    d.mu_smem = 0;
    if (dp.smemB + sizeof(double) * (size_t)L * 16 * d.MS <= 220 * 1024) {
        d.mu_smem = 1;
        dp.smemB += sizeof(double) * (size_t)L * 16 * d.MS;
    }
    if (dp.smemF > 227 * 1024 || dp.smemB > 227 * 1024) { dp.strip_ok = false; dp.strip_err = "dense strip kernels: shared-memory tile does not fit (N or K*(L+1) too large)"; }
    if (dp.strip_ok) {
        cudaError_t e = cudaFuncSetAttribute(dense_chain<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.smemF);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_chain<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.smemF);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.smemB);
        if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(e); return GRAPE_B200_ECUDA; }
    }
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) { err = "device does not support cooperative launch"; return GRAPE_B200_ECUDA; }
    dp.ready = true;
    return 0;
}

inline void dense_run_forward(DensePlan& dp, const DevP& p, cudaStream_t st, int64_t& launches) {
    DenseDev& d = dp.d;
    const size_t splane = (size_t)d.Np * d.Kp;
    cudaMemcpyAsync(d.cur, d.psi0, 2 * splane * sizeof(double), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(d.store, d.psi0, 2 * splane * sizeof(double), cudaMemcpyDeviceToDevice, st);
    DevP pp = p;
    void* args[] = {&pp, &d, &dp.kd};
    cudaLaunchCooperativeKernel((void*)dense_chain<false>, dim3(dp.gridF), dim3(DENSE_THREADS), args, dp.smemF, st);
    dense_tau<<<p.K, 256, 0, st>>>(p, d, dp.gridF);
    launches += 2;
}
inline void dense_run_backward(DensePlan& dp, const DevP& p, const cplx* chi_host, cudaStream_t st, int64_t& launches) {
    DenseDev& d = dp.d;
    const size_t bplane = (size_t)d.Np * d.Cb;
    cudaMemsetAsync(d.bcur, 0, 2 * bplane * sizeof(double), st);
    dense_boundary<<<p.K, 256, 0, st>>>(p, d, chi_host, dp.kd.on ? dp.kd.kcur : nullptr);
    DevP pp = p;
    void* args[] = {&pp, &d};
    cudaLaunchCooperativeKernel((void*)dense_backward, dim3(dp.gridB), dim3(DENSE_THREADS), args, dp.smemB, st);
    launches += 2;
    if (dp.kd.on) {
        void* cargs[] = {&pp, &d, &dp.kd};
            cudaLaunchCooperativeKernel((void*)dense_chain<true>, dim3(dp.gridF), dim3(DENSE_THREADS), cargs, dp.smemF, st);
        launches += 1;
    }
}
inline void dense_gather_final(DensePlan& dp, const DevP& p, cplx* out, cudaStream_t st, int64_t& launches) {
    const int cnt = p.K * p.N;
    dense_gather_final_k<<<(cnt + 255) / 256, 256, 0, st>>>(dp.d.cur, out, dp.d.Np, dp.d.Kp, p.N, p.K);
    launches++;
}
inline void dense_gather_states(DensePlan& dp, const DevP& p, int k, cplx* out, cudaStream_t st, int64_t& launches) {
    const int cnt = (p.NT + 1) * p.N;
    dense_gather_states_k<<<(cnt + 255) / 256, 256, 0, st>>>(dp.d.store, out, dp.d.Np, dp.d.Kp, p.N, p.NT, k);
    launches++;
}
