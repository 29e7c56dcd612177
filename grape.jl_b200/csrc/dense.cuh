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
#include "econ.cuh"
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
    int conc;          // concurrent forward / backward chains (dense_chain<2, NS>): the chi chain propagates tgt_k / ||tgt_k||
    double* kfac;      // [2][Kp] (re, im) ||tgt_k|| conj(c_k): factor of e_b in kry_combine instead of rho_k
    const double* tgtn;// [2][Np][Kp] planar tgt_k / ||tgt_k||: start of the concurrent chi chain
    unsigned* bars;    // [2 directions][Kp / 8 column groups][32] arrival counters of the split-phase barriers (one 128-byte
                       // line each; zeroed before every concurrent launch)
};

struct DenseDev {
    int Np, Kp, Cb, RT, Pf, Pb, MS, CcapF, CcapB, nD;
    const int* kry_ok; // block-recursion backward kernels return at once if *kry_ok (nullptr: always run)
    int mu_smem;   // backward strip kernel keeps its 8 rows of every mu_l^dagger in shared memory
    int nstrip;    // strip chain: Taylor terms per grid barrier (1; 2: strips of H_n, H_n^2; 3: + H_n^3; dense_dual_setup)
    int nP3;       // (L+1)(L+2)(L+3)/6 symmetrised triple products
    const double* PTf;  // [nP3][2][Np][Np]  sum over the distinct orderings of H_i H_j H_k, i <= j <= k (lexicographic)
    const double* PTa;
    int tma;       // strip prefetch through the TMA copy engine (cp.async.bulk + mbarrier) instead of per-thread cp.async
    int herm;      // every operator equals its adjoint (exact): backward chains read the forward generators
    int econ;      // Hermitian generators: Krylov-form chains sum the economised polynomial (c_econ) instead of the Taylor series
    int nP;        // (L+1)(L+2)/2 pair products
    const double* PPf;  // [nP][2][Np][Np]  H_i H_j + H_j H_i (i < j), H_i^2 (i == j), pair order (0,0),(0,1)..(0,L),(1,1)..(L,L)
    const double* PPa;  // the same products of the adjoints
    // pre-formed generators of the current pulses, written once per call by dense_preform (all steps in parallel) and
    // prefetched strip by strip by the multi-term chain:  pre[n][2 nstrip][Np][Np] = {Re H_n, Im H_n, Re H_n^2, Im H_n^2, ..}
    double* preF;       // forward (H_n); nullptr: the chain forms its strips itself
    double* preA;       // backward (H_n^dagger); == preF for Hermitian generators
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
    size_t smemF2;        // dense_chain<BWD, NS >= 2>: further operator strips
    bool ready;
    bool concurrent;      // dense_chain<2, NS> available for this handle (no state running cost, built-in chi, 2 RT <= SMs)
    bool dual_launched;   // the forward phase of the current call enqueued the concurrent chains
    DenseDev dD;          // copy of d with the column split of the concurrent launch (Pf, CcapF)
    int gridD;
    size_t smemD;
    bool strip_ok;        // the 8-row strip kernels of this file can run (N, K(L+1) small enough)
    std::string strip_err;
    DensePlan() : kry_smem(0), gridF(0), gridB(0), smemF(0), smemB(0), smemF2(0), ready(false), concurrent(false), dual_launched(false), gridD(0), smemD(0), strip_ok(true) { memset(&kd, 0, sizeof kd); memset(&dD, 0, sizeof dD); }
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

// One 8-column group, NS >= 2 operand strips: acc[q] += B_q X with every X fragment loaded once (B_q = the CTA's rows
// of H_n^(q+1): NS Taylor terms per grid barrier).  4 NS independent DMMA chains per warp (8 warps share the pipe, so
// a chain is revisited every >= 128 issue clocks, above the ~100 clk latency of a dependent DMMA).
// Bq[q] = strip q in shared memory (re plane, im plane at + 8 * bstride).
template <int NS, int UK>
GB_D void dense_mma_multi(const double* const (&Bq)[3], int bstride,
                          const double* __restrict__ Xre, const double* __restrict__ Xim, int ldx, int col0,
                          int kbeg, int kend, DAcc (&acc)[NS]) {
    const int lane = threadIdx.x & 31;
    const int lr = lane >> 2, lc = lane & 3;
    for (int kb = kbeg; kb < kend; kb += 4 * UK) {
        double are[UK], aim[UK];
#pragma unroll
        for (int u = 0; u < UK; ++u) {
            const int k0 = kb + 4 * u;
            if (k0 < kend) {
                const size_t off = (size_t)(k0 + lc) * ldx + col0 + lr;
                are[u] = __ldcg(&Xre[off]);
                aim[u] = __ldcg(&Xim[off]);
            }
        }
#pragma unroll
        for (int u = 0; u < UK; ++u) {
            const int k0 = kb + 4 * u;
            if (k0 < kend) {
                const int bo = lr * bstride + k0 + lc;
                double br[NS], bi[NS];
#pragma unroll
                for (int q = 0; q < NS; ++q) { br[q] = Bq[q][bo]; bi[q] = Bq[q][8 * bstride + bo]; }
#pragma unroll
                for (int q = 0; q < NS; ++q) dmma884(acc[q].p1, are[u], br[q]);
#pragma unroll
                for (int q = 0; q < NS; ++q) dmma884(acc[q].p2, aim[u], bi[q]);
#pragma unroll
                for (int q = 0; q < NS; ++q) dmma884(acc[q].q1, are[u], bi[q]);
#pragma unroll
                for (int q = 0; q < NS; ++q) dmma884(acc[q].q2, aim[u], br[q]);
            }
        }
    }
}

// Cross-warp (k-split) reduction of NS results of one 8-column group: afterwards thread t < 64 holds ALL NS complex
// results for row nrow = t / 8, column m = t % 8 (so that one thread adds the terms in a fixed order).
template <int NS>
GB_D void dense_reduce_all(double* __restrict__ red, DAcc (&acc)[NS], cplx (&res)[NS]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        double* rr = red + (size_t)((w * NS + q) * 2) * 64;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            rr[(2 * lc + e) * 8 + lr] = acc[q].p1[e] - acc[q].p2[e];
            rr[64 + (2 * lc + e) * 8 + lr] = acc[q].q1[e] + acc[q].q2[e];
        }
    }
    __syncthreads();
    if (threadIdx.x < 64) {
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            cplx r = mk(0.0, 0.0);
#pragma unroll
            for (int ww = 0; ww < DENSE_THREADS / 32; ++ww) {
                const double* rr = red + (size_t)((ww * NS + q) * 2) * 64;
                r.x += rr[threadIdx.x];
                r.y += rr[64 + threadIdx.x];
            }
            res[q] = r;
        }
    }
    __syncthreads();
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

// Taylor order / sub-steps of step n (block recursion), or -- econ -- the degree and weights of the economised polynomial
GB_D double dense_theta(const DevP& p, const DenseDev& d, int n, double dt) {
    double nrm = d.hnorm[0];
    for (int l = 0; l < p.L; ++l) {
        double a = p.eps[l * p.NT + n];
        if (p.shape) a *= p.shape[l * p.NT + n];
        nrm += fabs(a) * d.hnorm[1 + l];
    }
    return nrm * dt;
}
GB_D void dense_plan(const DevP& p, const DenseDev& d, int n, double dt, int& m, int& s) {
    vec_plan(dense_theta(p, d, n, dt), m, s);
}
GB_D void dense_plan(const DevP& p, const DenseDev& d, int n, double dt, int& m, int& s, const double*& gw, bool econ) {
    const double th = dense_theta(p, d, n, dt);
    gw = c_econ.ones;
    if (econ && th <= c_econ.theta[ECON_MAXM]) {
        m = 2;
        while (c_econ.theta[m] < th) ++m;
        s = 0;
        gw = c_econ.g[m];
        return;
    }
    vec_plan(th, m, s);
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

// rows r0..r0+7 of  sum_q coef[q] mats[q]  (planar matrices, coef in shared memory) into shared memory
GB_D void dense_form_comb(const double* __restrict__ mats, int nmat, const double* __restrict__ coef, int Np, int MS,
                          int r0, double* __restrict__ Hs_re, double* __restrict__ Hs_im) {
    const size_t plane = (size_t)Np * Np;
    constexpr int U = 8;
    const int tot = 8 * Np;
    const double* __restrict__ row = mats + (size_t)r0 * Np;
    for (int e0 = threadIdx.x; e0 < tot; e0 += U * DENSE_THREADS) {
        double hr[U], hi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) hr[u] = hi[u] = 0.0;
        for (int q = 0; q < nmat; ++q) {
            const double cq = coef[q];
            const double* __restrict__ Hq = row + (size_t)q * 2 * plane;
            double tr[U], ti[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * DENSE_THREADS;
                tr[u] = e < tot ? __ldg(&Hq[e]) : 0.0;
                ti[u] = e < tot ? __ldg(&Hq[plane + e]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                hr[u] = fma(cq, tr[u], hr[u]);
                hi[u] = fma(cq, ti[u], hi[u]);
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

// H_n, H_n^2 [and H_n^3] of EVERY time step of the current pulses (embarrassingly parallel, HBM-write bound): thread =
// one matrix element, blockIdx.y = chunk of time steps; the (L+1) + (L+1)(L+2)/2 [+ (L+1)(L+2)(L+3)/6] source values
// stay in registers over the chunk.  H_n^2 = sum_{i<=j} c_i c_j P_ij,  H_n^3 = sum_{i<=j<=k} c_i c_j c_k S_ijk.
template <int L, int NS>
__global__ void __launch_bounds__(256) dense_preform(DevP p, const double* __restrict__ Hall, const double* __restrict__ PP,
                                                     const double* __restrict__ PT, double* __restrict__ out, int Np, int chunk) {
    const size_t plane = (size_t)Np * Np;
    const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= plane) return;
    constexpr int PM = (L + 1) * (L + 2) / 2, TM = NS >= 3 ? (L + 1) * (L + 2) * (L + 3) / 6 : 1;
    double hr[L + 1], hi[L + 1], qr[PM], qi[PM], tr[TM], ti[TM];
#pragma unroll
    for (int q = 0; q <= L; ++q) { hr[q] = __ldg(&Hall[(size_t)q * 2 * plane + e]); hi[q] = __ldg(&Hall[(size_t)q * 2 * plane + plane + e]); }
#pragma unroll
    for (int q = 0; q < PM; ++q) { qr[q] = __ldg(&PP[(size_t)q * 2 * plane + e]); qi[q] = __ldg(&PP[(size_t)q * 2 * plane + plane + e]); }
    if (NS >= 3) {
#pragma unroll
        for (int q = 0; q < TM; ++q) { tr[q] = __ldg(&PT[(size_t)q * 2 * plane + e]); ti[q] = __ldg(&PT[(size_t)q * 2 * plane + plane + e]); }
    }
    const int n0 = blockIdx.y * chunk, n1 = min(p.NT, n0 + chunk);
    for (int n = n0; n < n1; ++n) {
        double c[L + 1];
        c[0] = 1.0;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            double a = p.eps[l * p.NT + n];
            if (p.shape) a *= p.shape[l * p.NT + n];
            c[1 + l] = a;
        }
        double ar = 0.0, ai = 0.0, br = 0.0, bi = 0.0, cr = 0.0, ci = 0.0;
#pragma unroll
        for (int q = 0; q <= L; ++q) { ar = fma(c[q], hr[q], ar); ai = fma(c[q], hi[q], ai); }
        int q = 0, q3 = 0;   // orders of dense_dual_setup: (0,0),(0,1)..(L,L) and (0,0,0),(0,0,1)..(L,L,L)
#pragma unroll
        for (int i = 0; i <= L; ++i)
#pragma unroll
            for (int j = i; j <= L; ++j, ++q) {
                const double cc = c[i] * c[j];
                br = fma(cc, qr[q], br);
                bi = fma(cc, qi[q], bi);
                if (NS >= 3) {
#pragma unroll
                    for (int k = j; k <= L; ++k, ++q3) {
                        const double c3 = cc * c[k];
                        cr = fma(c3, tr[q3], cr);
                        ci = fma(c3, ti[q3], ci);
                    }
                }
            }
        double* o = out + (size_t)n * (2 * NS) * plane + e;
        __stcs(&o[0], ar);
        __stcs(&o[plane], ai);
        __stcs(&o[2 * plane], br);
        __stcs(&o[3 * plane], bi);
        if (NS >= 3) {
            __stcs(&o[4 * plane], cr);
            __stcs(&o[5 * plane], ci);
        }
    }
}

// cp.async of the 2 NS 8-row strips {Re H, Im H, Re H^2, Im H^2, ..} of step n into shared memory (row stride MS):
// strip 0 at Hs (re, im), strips 1.. at HX
template <int NS>
GB_D void dense_prefetch_strips(const double* __restrict__ pre, int n, int Np, int MS, int r0,
                                double* __restrict__ Hs, double* __restrict__ HX) {
    const size_t plane = (size_t)Np * Np;
    const double* src = pre + (size_t)n * (2 * NS) * plane + (size_t)r0 * Np;
    const int segs = Np / 2;                 // 16-byte segments per row
    const int tot = 2 * NS * 8 * segs;
    for (int e = threadIdx.x; e < tot; e += DENSE_THREADS) {
        const int sg = e % segs, r = (e / segs) & 7, q = e / (8 * segs);   // q = plane index 0 .. 2 NS - 1
        double* dst = (q < 2 ? Hs + q * 8 * MS : HX + (q - 2) * 8 * MS) + r * MS + 2 * sg;
        cp_async16(dst, src + (size_t)q * plane + (size_t)r * Np + 2 * sg);
    }
    cp_async_commit();
}

// Split-phase barrier of one (direction, 8-column group): the Taylor terms of a column group only depend on the SAME
// group's previous term on all row tiles, so every group gets its own arrival counter.  A CTA that owns two groups
// arrives for group A, computes group B, and only then waits for A: the barrier round trip and the arrival skew of
// the other CTAs are hidden behind DMMA work instead of being paid per stage (cooperative_groups' grid.sync() of the
// whole grid cost ~40 % of the chain).  All CTAs are co-resident (cooperative launch), so spinning is safe; a wait
// gives up after ~10 s and flags the call instead of hanging the GPU.
GB_D unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
GB_D void split_arrive(unsigned* bar) {   // call after __syncthreads(), by one thread
    __threadfence();
    atomicAdd(bar, 1u);
}
GB_D void split_wait(const unsigned* bar, unsigned target, DevFlags* flags) {   // by one thread, then __syncthreads()
    if (ld_acquire_gpu_u32(bar) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu_u32(bar) < target) {
        if (clock64() - t0 > 20000000000ll) { flags->xchg_timeout = 2; break; }
    }
}

// The same strips through the TMA copy engine: thread 0 arms the mbarrier with the byte count and issues one bulk copy
// per (plane, row) -- 2 NS x 8 rows of Np doubles, 184 KB per step for C4 -- instead of 45 LDGSTS per thread.
template <int NS>
GB_D void dense_prefetch_strips_tma(const double* __restrict__ pre, int n, int Np, int MS, int r0,
                                    double* __restrict__ Hs, double* __restrict__ HX, unsigned long long* bar) {
    if (threadIdx.x == 0) {
        const size_t plane = (size_t)Np * Np;
        const double* src = pre + (size_t)n * (2 * NS) * plane + (size_t)r0 * Np;
        const unsigned row_bytes = (unsigned)Np * sizeof(double);
        fence_proxy_async();
        mbar_arrive_expect_tx(bar, 2 * NS * 8 * row_bytes);
#pragma unroll 1
        for (int q = 0; q < 2 * NS; ++q) {
            double* dq = q < 2 ? Hs + q * 8 * MS : HX + (q - 2) * 8 * MS;
#pragma unroll
            for (int r = 0; r < 8; ++r) bulk_g2s(dq + r * MS, src + (size_t)q * plane + (size_t)r * Np, row_bytes, bar);
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
// NS >= 2: further strips hold the same 8 rows of H_n^2 (and H_n^3); H_n^2 = sum_{i<=j} c_i c_j P_ij (c = (1, a_1 .. a_L), P_ij precomputed
// once per handle), and every grid barrier of a Krylov-form step separates NS Taylor terms,
//   T_{j+q} = (-+i dt)^q / ((j+1)..(j+q)) H^q T_j,   q = 1..NS,
// computed from the same operand fragments: ceil(m/NS) instead of m dependent stages per step.  With the pre-formed
// generators (DenseDev::preF / preA; required for NS = 3) the strips of step n+1 are fetched by cp.async behind the
// last stage of step n.
// MODE 0: forward sweep, MODE 1: chi chain, MODE 2: BOTH AT ONCE -- the first half of the grid runs the forward sweep
// upwards in n while the second half runs the chi chain downwards from n = NT - 1, sharing the grid barriers.  That is
// possible because chi_k(T) of J_T_sm / J_T_re / J_T_ss is c_k tgt_k with a SCALAR c_k that depends on the forward
// result (optimize.jl:845-855, docs/src/tutorial.md:399-405) and the backward propagation is linear: the chain
// propagates tgt_k / ||tgt_k||, and the factor ||tgt_k|| conj(c_k) (= rho_k times the phase of c_k) goes into the
// combination e_b of the contraction (kry_combine) once the forward sweep has produced tau.  Not applicable with a
// state running cost (its inhomogeneity needs Psi(t_{n-1}) while chi is propagated, optimize.jl:897-908) nor with
// a host chi.  A 2 N_T m / NS-stage latency chain becomes N_T m / NS stages of twice the width.
// skip_if_ok: return at once when the call qualifies for the Krylov form (MODE 2 has done / will do this sweep).
template <int MODE, int NS>
__global__ void __launch_bounds__(DENSE_THREADS, 1) dense_chain(DevP p, DenseDev d, KryDev kd, int skip_if_ok) {
    if (MODE >= 1 && !(*kd.ok)) return;   // uniform over the grid
    if (skip_if_ok && *kd.ok) return;
    extern __shared__ __align__(16) double dsm[];
    __shared__ __align__(8) unsigned long long strip_bar;   // completion barrier of the TMA strip copies
    const int Np = d.Np, Kp = d.Kp, MS = d.MS, NT = p.NT;
    const int half = MODE == 2 ? (int)(gridDim.x / 2) : 0;
    const bool BWD = MODE == 1 || (MODE == 2 && (int)blockIdx.x >= half);
    const int bid = (MODE == 2 && BWD) ? (int)blockIdx.x - half : (int)blockIdx.x;
    const int rt = bid / d.Pf, part = bid % d.Pf;
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
    double* coef = jb_s + Ccap + 8;                 // NS >= 2: pair coefficients c_i c_j of the current step (<= 64)
    double* HX = coef + 64;                         // NS >= 2: rows of H_n^2 [, H_n^3], 16 MS doubles each
    constexpr bool DUAL = NS >= 2;
    constexpr bool SPLIT = MODE == 2 && NS >= 2;   // per-(direction, column group) split-phase barriers instead of grid.sync()
    unsigned* const mybars = SPLIT ? kd.bars + (size_t)(BWD ? 1 : 0) * (Kp / 8) * 32 : nullptr;
    unsigned round = 0;                            // arrivals this CTA has made per column group
    auto stages_of = [&](int mm) { return DUAL ? (mm + NS - 1) / NS : mm; };   // grid barriers of a Krylov-form step
    const double* const Bq[3] = {Hs_re, HX, HX + 16 * MS};
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

    const double* pre = DUAL ? (BWD ? d.preA : d.preF) : nullptr;
    const bool tma = DUAL && pre && d.tma;
    unsigned strip_phase = 0;
    if (tma) {
        if (threadIdx.x == 0) { mbar_init(&strip_bar, 1); fence_mbar_init(); }
        __syncthreads();
        dense_prefetch_strips_tma<NS>(pre, BWD ? NT - 1 : 0, Np, MS, r0, Hs_re, HX, &strip_bar);
    } else if (DUAL && pre) dense_prefetch_strips<NS>(pre, BWD ? NT - 1 : 0, Np, MS, r0, Hs_re, HX);
    for (int it = 0; it < NT; ++it) {
        const int n = BWD ? NT - 1 - it : it;
        const double dt = p.tlist[n + 1] - p.tlist[n];
        if (!BWD && gb) {
            const double wgt = n == 0 ? 0.5 * (p.tlist[1] - p.tlist[0]) : 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]);
            gb_point(wgt);
        }
        if (tma) { mbar_wait(&strip_bar, strip_phase); strip_phase ^= 1; }   // strips of this step, issued one step ahead
        else if (DUAL && pre) cp_async_wait<0>();   // strips of this step: issued behind the previous step's last stage
        else dense_form_H(p, Hall, Np, MS, r0, n, Hs_re, Hs_im);
        int m, s;
        const double* gw;   // weights of the Taylor terms in the new state (1, or the economised polynomial's)
        dense_plan(p, d, n, dt, m, s, gw, d.econ && kd.on);
        if (!BWD && p.grad_method != 0 && p.taylor_check && m > p.taylor_max_order && bid == 0 && threadIdx.x == 0)
            p.flags->taylor_fail = 1;
        const bool kry = kd.on && s == 0 && m <= kd.MT;
        double* slots = kry ? terms + (size_t)n * kd.MT * 2 * splane : nullptr;
        const bool dual = DUAL && kry && m >= 2;
        if (DUAL && dual && !pre) {   // uniform over the grid
            // c = (1, a_1 .. a_L); pair (i, j), i <= j, at index i (L+1) - i (i-1)/2 + (j - i)
            if (threadIdx.x < d.nP) {
                int i = 0, q = threadIdx.x;
                while (q >= p.L + 1 - i) { q -= p.L + 1 - i; ++i; }
                const int jj = i + q;
                auto cf = [&](int t) {
                    if (t == 0) return 1.0;
                    double a = p.eps[(t - 1) * p.NT + n];
                    if (p.shape) a *= p.shape[(t - 1) * p.NT + n];
                    return a;
                };
                coef[threadIdx.x] = cf(i) * cf(jj);
            }
            __syncthreads();
            dense_form_comb(BWD ? d.PPa : d.PPf, d.nP, coef, Np, MS, r0, HX, HX + 8 * MS);
        }
        __syncthreads();
        const int nsub = 1 << s;
        const double dts = dt / nsub;
        for (int sub = 0; sub < nsub; ++sub) {
            if (DUAL && dual) {
                // NS Taylor terms per stage (nsub == 1 here): T_{j+q} = f_q H^q T_j, f_q = (-+i dts)^q / ((j+1)..(j+q))
                for (int j = 0; j < m; j += NS) {
                    const int nt = min(NS, m - j);   // terms of this stage
                    const double* src = j == 0 ? state : slots + (size_t)j * 2 * splane;
                    double xq[NS];
                    {
                        double x = 1.0;
#pragma unroll
                        for (int q = 0; q < NS; ++q) { x *= dts / (j + q + 1); xq[q] = x; }
                    }
                    // f_q res: q = 0: (-+i) x res; q = 1: -x res; q = 2: (+-i) x res   (forward: -i, backward: +i)
                    auto term = [&](int q, double x, cplx r, double& tr, double& ti) {
                        if (q == 0) { tr = BWD ? -x * r.y : x * r.y; ti = BWD ? x * r.x : -x * r.x; }
                        else if (q == 1) { tr = -x * r.x; ti = -x * r.y; }
                        else { tr = BWD ? x * r.y : -x * r.y; ti = BWD ? -x * r.x : x * r.x; }
                    };
                    const int pgstep = nt == NS ? 1 : DENSE_CGP;   // full stage: one 8-column group per pass, all NS strips
                    for (int pg = cg0; pg < cg1; pg += pgstep) {
                        const int ng = min(pgstep, cg1 - pg);
                        const int g = threadIdx.x >> 6, idx = threadIdx.x & 63, nrow = idx >> 3, mc = idx & 7;
                        if (SPLIT) {   // the operand columns of these groups are complete on every row tile
                            if (threadIdx.x == 0)
                                for (int gg = 0; gg < ng; ++gg) split_wait(mybars + (size_t)(pg + gg) * 32, round * (unsigned)d.RT, p.flags);
                            __syncthreads();
                        }
                        if (nt == NS) {
                            DAcc acc[NS];
#pragma unroll
                            for (int q = 0; q < NS; ++q) acc[q].zero();
                            dense_mma_multi<NS, 8>(Bq, MS, src, src + splane, Kp, pg * 8, kbeg, kend, acc);
                            cplx res[NS];
                            dense_reduce_all<NS>(red, acc, res);   // threads 0..63 hold all NS results
                            if (threadIdx.x < 64) {
                                const int cglob = pg * 8 + mc;
                                const size_t off = (size_t)(r0 + nrow) * Kp + cglob;
                                double ar = acc_re[nrow * Ccap + cglob - cbeg], ai = acc_im[nrow * Ccap + cglob - cbeg];
#pragma unroll
                                for (int q = 0; q < NS; ++q) {   // fixed summation order: term j+1, j+2, ..
                                    double tr, ti;
                                    term(q, xq[q], res[q], tr, ti);
                                    if (j + q + 1 < m) {
                                        double* dst = slots + (size_t)(j + q + 1) * 2 * splane;
                                        dst[off] = tr;
                                        dst[splane + off] = ti;
                                    }
                                    const double gq = gw[j + q + 1];
                                    ar = fma(gq, tr, ar);
                                    ai = fma(gq, ti, ai);
                                }
                                acc_re[nrow * Ccap + cglob - cbeg] = ar;
                                acc_im[nrow * Ccap + cglob - cbeg] = ai;
                            }
                            if (SPLIT && j + NS < m) {   // this group's new terms are written: arrive, go on with the next group
                                __syncthreads();
                                if (threadIdx.x == 0) split_arrive(mybars + (size_t)pg * 32);
                            }
                        } else {
                            // several column groups per CTA (or a short last stage): one pass per strip
                            for (int q = 0; q < nt; ++q) {
                                int col0[DENSE_CGP];
                                DAcc acc[DENSE_CGP];
#pragma unroll
                                for (int gg = 0; gg < DENSE_CGP; ++gg) { col0[gg] = (pg + gg) * 8; acc[gg].zero(); }
                                dense_mma_slice<DENSE_CGP, true, 2>(Bq[q], Bq[q] + 8 * MS, MS, 1.0, src, src + splane, Kp, col0, ng,
                                                                    kbeg, kend, acc);
                                const cplx res = dense_reduce<DENSE_CGP>(red, acc, ng);
                                if (g < ng) {
                                    const int cglob = (pg + g) * 8 + mc;
                                    const size_t off = (size_t)(r0 + nrow) * Kp + cglob;
                                    double tr, ti, x = 1.0;
                                    for (int t = 0; t <= q; ++t) x *= dts / (j + t + 1);
                                    term(q, x, res, tr, ti);
                                    if (j + q + 1 < m) {
                                        double* dst = slots + (size_t)(j + q + 1) * 2 * splane;
                                        dst[off] = tr;
                                        dst[splane + off] = ti;
                                    }
                                    const double gq = gw[j + q + 1];
                                    acc_re[nrow * Ccap + cglob - cbeg] = fma(gq, tr, acc_re[nrow * Ccap + cglob - cbeg]);
                                    acc_im[nrow * Ccap + cglob - cbeg] = fma(gq, ti, acc_im[nrow * Ccap + cglob - cbeg]);
                                }
                            }
                        }
                    }
                    if (SPLIT) { if (j + NS < m) ++round; }
                    else if (j + NS < m) grid.sync();
                }
            } else
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
                        acc_re[nrow * Ccap + cglob - cbeg] = fma(gw[j], tr, acc_re[nrow * Ccap + cglob - cbeg]);
                        acc_im[nrow * Ccap + cglob - cbeg] = fma(gw[j], ti, acc_im[nrow * Ccap + cglob - cbeg]);
                    }
                }
                if (j < m) grid.sync();
            }
            __syncthreads();
            const bool last = sub == nsub - 1;
            if (DUAL && pre && last && it + 1 < NT) {   // every warp is done with this step's strips: fetch the next ones
                if (tma) dense_prefetch_strips_tma<NS>(pre, BWD ? n - 1 : n + 1, Np, MS, r0, Hs_re, HX, &strip_bar);
                else dense_prefetch_strips<NS>(pre, BWD ? n - 1 : n + 1, Np, MS, r0, Hs_re, HX);   // behind the epilogue and the barrier
            }
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
            if (SPLIT && stages_of(m) == 1) {
                // a one-stage step has no barrier between reading `state` (the operand of its only stage) and overwriting
                // it below: one extra round so that every row tile is done reading
                __syncthreads();
                if (threadIdx.x == 0)
                    for (int pg = cg0; pg < cg1; ++pg) split_arrive(mybars + (size_t)pg * 32);
                ++round;
                if (threadIdx.x == 0)
                    for (int pg = cg0; pg < cg1; ++pg) split_wait(mybars + (size_t)pg * 32, round * (unsigned)d.RT, p.flags);
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
            if (SPLIT) {   // the new state of every column group of this CTA is written
                __syncthreads();
                if (threadIdx.x == 0)
                    for (int pg = cg0; pg < cg1; ++pg) split_arrive(mybars + (size_t)pg * 32);
                ++round;
            } else grid.sync();
        }
        if (MODE == 2 && !SPLIT) {   // both directions make the same number of grid barriers per iteration (uniform per role)
            const int mo = kd.m_n[BWD ? it : NT - 1 - it];
            for (int x = stages_of(m); x < stages_of(mo); ++x) grid.sync();
        }
    }
    if (!BWD && gb) {
        gb_point(0.5 * (p.tlist[NT] - p.tlist[NT - 1]));
        for (int c = threadIdx.x; c < ncols; c += DENSE_THREADS)
            d.jbpart[(size_t)bid * Kp + cbeg + c] = jb_s[c];
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
                                const double sl = p.dshape ? p.dshape[l * NT + n] : 1.0;
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
__global__ void __launch_bounds__(256) dense_boundary(DevP p, DenseDev d, const cplx* __restrict__ chi_host, double* __restrict__ kcur,
                                                      double* __restrict__ kfac) {
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
        if (kfac) {   // concurrent chains: chi_k = (c_k / |c_k|) x (propagated tgt_k / ||tgt_k||), rho_k = |c_k| ||tgt_k||
            const double ac = sqrt(c.x * c.x + c.y * c.y);
            kfac[k] = ac > 0.0 ? rho * c.x / ac : 0.0;
            kfac[d.Kp + k] = ac > 0.0 ? -rho * c.y / ac : 0.0;
        }
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
    d.herm = (hf == ha) ? 1 : 0;
    {
        // economised polynomial instead of the Taylor series on the Krylov-form schedule (GRAPE_B200_ECON=0: Taylor)
        const char* e = getenv("GRAPE_B200_ECON");
        d.econ = d.herm && !(e && atoi(e) == 0);
        static bool uploaded[64];   // the table does not depend on the problem: once per device
        if (!(dev >= 0 && dev < 64 && uploaded[dev])) {
            if (cudaMemcpyToSymbol(c_econ, &econ_table(), sizeof(EconTab)) != cudaSuccess) { err = "cudaMemcpyToSymbol failed (economised-polynomial table)"; return GRAPE_B200_ECUDA; }
            if (dev >= 0 && dev < 64) uploaded[dev] = true;
        }
    }
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
        cudaError_t e = cudaFuncSetAttribute(dense_chain<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.smemF);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_chain<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.smemF);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.smemB);
        if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(e); return GRAPE_B200_ECUDA; }
    }
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) { err = "device does not support cooperative launch"; return GRAPE_B200_ECUDA; }
    dp.ready = true;
    return 0;
}

// C (+)= A B (+ B A if sym): planar complex Np x Np, one thread per output element (set-up only)
__global__ void __launch_bounds__(256) dense_pair_product(const double* __restrict__ A, const double* __restrict__ B,
                                                          double* __restrict__ C, int Np, int sym, int accum) {
    const size_t plane = (size_t)Np * Np;
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= Np || j >= Np) return;
    double cr = accum ? C[(size_t)i * Np + j] : 0.0, ci = accum ? C[plane + (size_t)i * Np + j] : 0.0;
    for (int k = 0; k < Np; ++k) {
        const double ar = A[(size_t)i * Np + k], ai = A[plane + (size_t)i * Np + k];
        const double br = B[(size_t)k * Np + j], bi = B[plane + (size_t)k * Np + j];
        cr = fma(ar, br, cr); cr = fma(-ai, bi, cr);
        ci = fma(ar, bi, ci); ci = fma(ai, br, ci);
        if (sym) {
            const double xr = B[(size_t)i * Np + k], xi = B[plane + (size_t)i * Np + k];
            const double yr = A[(size_t)k * Np + j], yi = A[plane + (size_t)k * Np + j];
            cr = fma(xr, yr, cr); cr = fma(-xi, yi, cr);
            ci = fma(xr, yi, ci); ci = fma(xi, yr, ci);
        }
    }
    C[(size_t)i * Np + j] = cr;
    C[plane + (size_t)i * Np + j] = ci;
}

// Several Taylor terms per grid barrier in the strip chains (dense_chain<BWD, NS>, NS = 2 or 3): needs the
// Krylov-form term storage, at most 3 controls and room for NS operator strips in shared memory; NS = 3 also needs
// the pre-formed generators (NT x 6 planes of device memory; x 2 for non-Hermitian generators).
// GRAPE_B200_DENSE_TERMS=1|2|3 caps NS, GRAPE_B200_DENSE_PREFORM=0 makes the chain form its strips itself (NS <= 2).
// Called after kry_setup.
inline int dense_dual_setup(DensePlan& dp, const DevP& p, bool tiled_chains, std::vector<void*>& allocs, std::string& err) {
    DenseDev& d = dp.d;
    d.nstrip = 1;
    d.preF = d.preA = nullptr;
    int want = 3;
    if (const char* env = getenv("GRAPE_B200_DENSE_TERMS")) want = atoi(env);
    if (const char* env = getenv("GRAPE_B200_DENSE_DUAL")) { if (atoi(env) == 0) want = 1; }
    if (want < 2 || !dp.kd.on || p.L > 3) return 0;
    if (!tiled_chains && !dp.strip_ok) return 0;
    const char* envp = getenv("GRAPE_B200_DENSE_PREFORM");
    const bool allow_pre = !(envp && atoi(envp) == 0);
    const size_t hplane = (size_t)d.Np * d.Np;
    auto smem_of = [&](int ns) { return dp.smemF + sizeof(double) * (64 + (size_t)(ns - 1) * 16 * d.MS); };
    int ns = std::min(want, 3);
    if (tiled_chains) {
        if (!allow_pre) return 0;   // the tiled multi-term chain (dense2_chain_multi) reads pre-formed generators only
    } else {
        if (ns == 3 && (smem_of(3) > 227 * 1024 || !allow_pre)) ns = 2;
        if (smem_of(2) > 227 * 1024) return 0;
    }
    // pre-formed generators: NT x 2 ns planes
    auto try_pre = [&](int nsx) -> bool {
        const size_t bytes = sizeof(double) * (size_t)p.NT * 2 * nsx * hplane;
        size_t freeB = 0, totB = 0;
        if (cudaMemGetInfo(&freeB, &totB) != cudaSuccess) { cudaGetLastError(); return false; }
        if ((double)bytes * (d.herm ? 1 : 2) > 0.6 * (double)freeB) return false;
        double* a = nullptr;
        double* b = nullptr;
        if (cudaMalloc((void**)&a, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
        if (d.herm) b = a;
        else if (cudaMalloc((void**)&b, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(a); return false; }
        allocs.push_back(a);
        if (b != a) allocs.push_back(b);
        d.preF = a; d.preA = b;
        return true;
    };
    if (allow_pre) {
        if (ns == 3 && !try_pre(3)) ns = 2;
        if (ns == 2 && !d.preF) try_pre(2);
    }
    if (tiled_chains && !d.preF) return 0;
    // operator products (once per handle)
    const int nP = (p.L + 1) * (p.L + 2) / 2, nP3 = (p.L + 1) * (p.L + 2) * (p.L + 3) / 6;
    auto dalloc = [&](double** q, size_t n) -> bool {
        if (cudaMalloc((void**)q, sizeof(double) * n) != cudaSuccess) { cudaGetLastError(); return false; }
        allocs.push_back(*q);
        return true;
    };
    double *pf = nullptr, *pa = nullptr, *tf = nullptr, *ta = nullptr, *tmp = nullptr;
    // Hermitian generators: the adjoints are the operators themselves, one set of products serves both sweeps
    if (!dalloc(&pf, (size_t)nP * 2 * hplane)) return 0;
    if (d.herm) pa = pf;
    else if (!dalloc(&pa, (size_t)nP * 2 * hplane)) return 0;
    if (ns == 3) {
        if (!dalloc(&tf, (size_t)nP3 * 2 * hplane) || !dalloc(&tmp, 2 * hplane)) return 0;
        if (d.herm) ta = tf;
        else if (!dalloc(&ta, (size_t)nP3 * 2 * hplane)) return 0;
    }
    dim3 grid((d.Np + 31) / 32, (d.Np + 7) / 8);
    for (int side = 0; side < (d.herm ? 1 : 2); ++side) {
        const double* Hm = side ? d.Ha : d.Hf;
        double* P = side ? pa : pf;
        double* T = side ? ta : tf;
        auto M = [&](int i) { return Hm + (size_t)i * 2 * hplane; };
        int q = 0, q3 = 0;
        for (int i = 0; i <= p.L; ++i)
            for (int j = i; j <= p.L; ++j, ++q) {
                dense_pair_product<<<grid, 256>>>(M(i), M(j), P + (size_t)q * 2 * hplane, d.Np, i != j, 0);
                if (ns < 3) continue;
                for (int k = j; k <= p.L; ++k, ++q3) {
                    // sum over the distinct orderings (a, b, c) of (i, j, k) of H_a (H_b H_c)
                    int perm[6][3] = {{i, j, k}, {i, k, j}, {j, i, k}, {j, k, i}, {k, i, j}, {k, j, i}};
                    int done = 0;
                    for (int t = 0; t < 6; ++t) {
                        bool dup = false;
                        for (int u = 0; u < t; ++u)
                            if (perm[u][0] == perm[t][0] && perm[u][1] == perm[t][1] && perm[u][2] == perm[t][2]) dup = true;
                        if (dup) continue;
                        dense_pair_product<<<grid, 256>>>(M(perm[t][1]), M(perm[t][2]), tmp, d.Np, 0, 0);
                        dense_pair_product<<<grid, 256>>>(M(perm[t][0]), tmp, T + (size_t)q3 * 2 * hplane, d.Np, 0, done ? 1 : 0);
                        done = 1;
                    }
                }
            }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { err = std::string("operator products failed: ") + cudaGetErrorString(cudaGetLastError()); return GRAPE_B200_ECUDA; }
    if (tiled_chains) {   // kernel attributes: dense2_multi_setup
        d.PPf = pf; d.PPa = pa; d.nP = nP; d.PTf = tf; d.PTa = ta; d.nP3 = nP3; d.nstrip = ns;
        return 0;
    }
    const size_t smem = smem_of(ns);
    cudaError_t e = ns == 3 ? cudaFuncSetAttribute(dense_chain<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(dense_chain<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
        e = ns == 3 ? cudaFuncSetAttribute(dense_chain<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                    : cudaFuncSetAttribute(dense_chain<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); d.preF = d.preA = nullptr; return 0; }
    d.PPf = pf; d.PPa = pa; d.nP = nP; d.PTf = tf; d.PTa = ta; d.nP3 = nP3; d.nstrip = ns;
    dp.smemF2 = smem;
    d.tma = d.preF && !(getenv("GRAPE_B200_DENSE_TMA") && atoi(getenv("GRAPE_B200_DENSE_TMA")) == 0) ? 1 : 0;
    return 0;
}

// Concurrent forward / backward chains (dense_chain<2, NS>): 2 x RT x Pf2 CTAs in one cooperative grid.
// Needs the Krylov form (term stores), a built-in functional (chi_k = c_k tgt_k), no state running cost, the strip
// kernels, and room for both directions on the device.  GRAPE_B200_DENSE_CONCURRENT=0 disables.  Called after
// dense_dual_setup; `tgt_host` = the ABI's [K][N] complex target states.
inline int dense_concurrent_setup(DensePlan& dp, const DevP& p, const double* tgt_host, bool tiled_chains,
                                  std::vector<void*>& allocs, std::string& err) {
    DenseDev& d = dp.d;
    dp.concurrent = false;
    dp.kd.conc = 0;
    if (const char* env = getenv("GRAPE_B200_DENSE_CONCURRENT")) { if (atoi(env) == 0) return 0; }
    if (!dp.kd.on || tiled_chains || !dp.strip_ok || p.gb_kind != 0 || p.functional == GRAPE_B200_JT_HOST) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (2 * d.RT > sms) return 0;
    DenseDev dd = d;
    dd.Pf = std::max(1, std::min(std::max(1, sms / (2 * d.RT)), d.Kp / 8));
    dd.CcapF = ((d.Kp / 8 + dd.Pf - 1) / dd.Pf) * 8;
    const size_t redB = (size_t)(DENSE_THREADS / 32) * DENSE_CGP * 128;
    size_t smem = sizeof(double) * (16 * (size_t)d.MS + 16 * (size_t)dd.CcapF + redB + dd.CcapF + 8);
    if (d.nstrip >= 2) smem += sizeof(double) * (64 + (size_t)(d.nstrip - 1) * 16 * d.MS);
    if (smem > 227 * 1024) return 0;
    cudaError_t e = d.nstrip == 3 ? cudaFuncSetAttribute(dense_chain<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                  : d.nstrip == 2 ? cudaFuncSetAttribute(dense_chain<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                  : cudaFuncSetAttribute(dense_chain<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    // normalised targets (static) and the per-call factors
    const size_t splane = (size_t)d.Np * d.Kp;
    std::vector<double> tn(2 * splane, 0.0);
    for (int k = 0; k < p.K; ++k) {
        double nn = 0.0;
        for (int i = 0; i < p.N; ++i) {
            const double re = tgt_host[2 * ((size_t)k * p.N + i)], im = tgt_host[2 * ((size_t)k * p.N + i) + 1];
            nn += re * re + im * im;
        }
        nn = std::sqrt(nn);
        if (!(nn > 0.0)) return 0;   // a trajectory without target state: chi_k(T) = 0, the sequential path reports it
        for (int i = 0; i < p.N; ++i) {
            tn[(size_t)i * d.Kp + k] = tgt_host[2 * ((size_t)k * p.N + i)] / nn;
            tn[splane + (size_t)i * d.Kp + k] = tgt_host[2 * ((size_t)k * p.N + i) + 1] / nn;
        }
    }
    void* q = nullptr;
    if (cudaMalloc(&q, tn.size() * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return 0; }
    allocs.push_back(q);
    if (cudaMemcpy(q, tn.data(), tn.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) { err = "cudaMemcpy failed"; return GRAPE_B200_ECUDA; }
    dp.kd.tgtn = static_cast<const double*>(q);
    if (cudaMalloc(&q, 2 * (size_t)d.Kp * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return 0; }
    allocs.push_back(q);
    cudaMemset(q, 0, 2 * (size_t)d.Kp * sizeof(double));
    dp.kd.kfac = static_cast<double*>(q);
    if (cudaMalloc(&q, 2 * (size_t)(d.Kp / 8) * 32 * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return 0; }
    allocs.push_back(q);
    dp.kd.bars = static_cast<unsigned*>(q);
    dp.dD = dd;
    dp.gridD = 2 * d.RT * dd.Pf;
    dp.smemD = smem;
    dp.kd.conc = 1;
    dp.concurrent = true;
    return 0;
}

template <int L>
inline void dense_preform_launch(const DenseDev& d, const DevP& p, bool adjoint, dim3 grid, int chunk, cudaStream_t st) {
    const double* Hs_ = adjoint ? d.Ha : d.Hf;
    const double* PP_ = adjoint ? d.PPa : d.PPf;
    const double* PT_ = adjoint ? d.PTa : d.PTf;
    double* out = adjoint ? d.preA : d.preF;
    if (d.nstrip == 3) dense_preform<L, 3><<<grid, 256, 0, st>>>(p, Hs_, PP_, PT_, out, d.Np, chunk);
    else dense_preform<L, 2><<<grid, 256, 0, st>>>(p, Hs_, PP_, PT_, out, d.Np, chunk);
}
inline void dense_run_preform(DensePlan& dp, const DevP& p, bool adjoint, cudaStream_t st, int64_t& launches) {
    DenseDev& d = dp.d;
    const int chunk = 32;
    dim3 grid((unsigned)(((size_t)d.Np * d.Np + 255) / 256), (unsigned)((p.NT + chunk - 1) / chunk));
    if (p.L == 1) dense_preform_launch<1>(d, p, adjoint, grid, chunk, st);
    else if (p.L == 2) dense_preform_launch<2>(d, p, adjoint, grid, chunk, st);
    else dense_preform_launch<3>(d, p, adjoint, grid, chunk, st);
    launches++;
}
template <int MODE>
inline void dense_chain_launch(DensePlan& dp, void** args, cudaStream_t st) {
    const DenseDev& d = dp.d;
    const int grid = MODE == 2 ? dp.gridD : dp.gridF;
    const size_t sm1 = MODE == 2 ? dp.smemD : dp.smemF, sm2 = MODE == 2 ? dp.smemD : dp.smemF2;
    if (d.nstrip == 3) cudaLaunchCooperativeKernel((void*)dense_chain<MODE, 3>, dim3(grid), dim3(DENSE_THREADS), args, sm2, st);
    else if (d.nstrip == 2) cudaLaunchCooperativeKernel((void*)dense_chain<MODE, 2>, dim3(grid), dim3(DENSE_THREADS), args, sm2, st);
    else cudaLaunchCooperativeKernel((void*)dense_chain<MODE, 1>, dim3(grid), dim3(DENSE_THREADS), args, sm1, st);
}

inline void dense_run_forward(DensePlan& dp, const DevP& p, cudaStream_t st, int64_t& launches, bool with_backward = false) {
    DenseDev& d = dp.d;
    const size_t splane = (size_t)d.Np * d.Kp;
    cudaMemcpyAsync(d.cur, d.psi0, 2 * splane * sizeof(double), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(d.store, d.psi0, 2 * splane * sizeof(double), cudaMemcpyDeviceToDevice, st);
    DevP pp = p;
    int skip = 0;
    if (d.nstrip > 1 && d.preF) dense_run_preform(dp, p, false, st, launches);
    dp.dual_launched = dp.concurrent && with_backward;
    if (dp.dual_launched) {
        // a gradient evaluation: both chains at once (the kernel returns at once unless the call qualifies for the
        // Krylov form; then the forward-only kernel below does the sweep as before)
        if (d.nstrip > 1 && d.preA && d.preA != d.preF) dense_run_preform(dp, p, true, st, launches);
        cudaMemcpyAsync(dp.kd.kcur, dp.kd.tgtn, 2 * splane * sizeof(double), cudaMemcpyDeviceToDevice, st);
        cudaMemsetAsync(dp.kd.bars, 0, 2 * (size_t)(d.Kp / 8) * 32 * sizeof(unsigned), st);
        void* dargs[] = {&pp, &dp.dD, &dp.kd, &skip};
        dense_chain_launch<2>(dp, dargs, st);
        launches++;
        skip = 1;
    }
    void* args[] = {&pp, &d, &dp.kd, &skip};
    dense_chain_launch<0>(dp, args, st);
    dense_tau<<<p.K, 256, 0, st>>>(p, d, dp.gridF);
    launches += 2;
}
inline void dense_run_backward(DensePlan& dp, const DevP& p, const cplx* chi_host, cudaStream_t st, int64_t& launches) {
    DenseDev& d = dp.d;
    const size_t bplane = (size_t)d.Np * d.Cb;
    const bool dual = dp.dual_launched && !chi_host;
    cudaMemsetAsync(d.bcur, 0, 2 * bplane * sizeof(double), st);
    // after concurrent chains kcur holds the propagated chi: the boundary kernel must not overwrite the term stores' input
    dense_boundary<<<p.K, 256, 0, st>>>(p, d, chi_host, (dp.kd.on && !dual) ? dp.kd.kcur : nullptr, dual ? dp.kd.kfac : nullptr);
    DevP pp = p;
    void* args[] = {&pp, &d};
    cudaLaunchCooperativeKernel((void*)dense_backward, dim3(dp.gridB), dim3(DENSE_THREADS), args, dp.smemB, st);
    launches += 2;
    if (dp.kd.on) {
        int skip = dual ? 1 : 0;
        void* cargs[] = {&pp, &d, &dp.kd, &skip};
        if (!dual && d.nstrip > 1 && d.preA && d.preA != d.preF) dense_run_preform(dp, p, true, st, launches);
        dense_chain_launch<1>(dp, cargs, st);
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
