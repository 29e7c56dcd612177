// dense2.cuh -- tiled FP64 tensor-core (DMMA m8n8k4) kernels of the large-N path.
//
// Same algorithm and HBM layout as dense.cuh (polynomial apply of exp(-i H_n dt) on the block of
// K states going forward, on the GradGenerator block [chi'_1 .. chi'_L, chi] of K(L+1) states going
// backward; persistent cooperative kernel, one grid barrier per Taylor term), but every term is a
// shared-memory tiled complex GEMM instead of an 8-row strip that re-reads the whole operand block
// from L2:
//   * CTA tile = TM(32) rows x TN(16) trajectories x ALL (L+1) blocks. One staged chunk of
//     {H_n^dagger, mu_1^dagger .. mu_L^dagger} rows and {chi'_1 .. chi'_L, chi} columns feeds the
//     2L+1 products of the GradGenerator block structure (docs/src/background.md:467-477):
//         out_l  = H^dagger chi'_l + s_l mu_l^dagger chi     (l < L),     out_L = H^dagger chi
//     so the chi tile and the H^dagger tile are loaded once and used L+1 times.
//   * 3-stage cp.async (LDGSTS) pipeline over the contraction index, padded planar tiles
//     (row stride = 20 doubles: conflict-free 64-bit fragment loads for both operands).
//   * 8 warps = 4 (row) x 2 (column) 8x8 DMMA tiles; complex product = 4 real DMMAs on split re/im.
//   * H_n (forward) / H_n^dagger (backward) is formed once per time step into a ping-pong buffer by
//     all CTAs (overlapped with the last term of the previous step).
//   * The running Taylor sum lives in a second state buffer (`nxt`) that only the owning thread
//     touches; cur/nxt swap at the end of a (sub)step.
#pragma once
#include "dense.cuh"

constexpr int D2_TM = 32, D2_TN = 16, D2_ST = 3;
constexpr int D2_KC_B = 16;   // contraction chunk of the backward (L+1 operand pairs per stage)
constexpr int D2_KC_F = 32;   // ... of the single-operand GEMMs (forward sweep, D Psi)
constexpr int D2_BS = D2_TN + 4;
constexpr int D2_THREADS = 256;
constexpr int D2_LBMAX = 5;   // L + 1 <= 5

struct Dense2Dev {
    int tilesR, tilesC, ntiles;
    double* Hn[2];    // formed generator, planar [2][Np][Np]
    double* nxt;      // planar [2][Np][Cb]
    double* cur2;     // forward ping-pong partner of DenseDev::cur, planar [2][Np][Kp]
};

struct D2Acc { double re[2], im[2]; };

__host__ __device__ constexpr size_t d2_stage_doubles(int LB, int KC) { return (size_t)LB * 2 * (D2_TM * (KC + 4) + KC * D2_BS); }

template <int LB, int D2_KC>
GB_D void d2_issue(double* __restrict__ stage, const double* const (&A)[LB], size_t aplane, int lda,
                   const double* const (&B)[LB], size_t bplane, int ldb, int k0) {
    constexpr int D2_AS = D2_KC + 4;
    constexpr int HS = D2_KC / 2;                       // 16-byte segments per A row
    constexpr int ASEG = 2 * D2_TM * HS;                // per matrix
    constexpr int TS = D2_TN / 2;
    constexpr int BSEG = 2 * D2_KC * TS;
    double* bst = stage + LB * 2 * D2_TM * D2_AS;
#pragma unroll
    for (int q = 0; q < LB; ++q) {
#pragma unroll
        for (int e0 = 0; e0 < ASEG; e0 += D2_THREADS) {
            const int e = e0 + threadIdx.x;
            const int s = e % HS, r = (e / HS) % D2_TM, pl = e / (HS * D2_TM);
            if (ASEG % D2_THREADS == 0 || e < ASEG)
                cp_async16(stage + ((q * 2 + pl) * D2_TM + r) * D2_AS + 2 * s,
                           A[q] + pl * aplane + (size_t)r * lda + k0 + 2 * s);
        }
#pragma unroll
        for (int e0 = 0; e0 < BSEG; e0 += D2_THREADS) {
            const int e = e0 + threadIdx.x;
            const int s = e % TS, kr = (e / TS) % D2_KC, pl = e / (TS * D2_KC);
            if (BSEG % D2_THREADS == 0 || e < BSEG)
                cp_async16(bst + ((q * 2 + pl) * D2_KC + kr) * D2_BS + 2 * s,
                           B[q] + pl * bplane + (size_t)(k0 + kr) * ldb + 2 * s);
        }
    }
}

template <int LB, bool MU, int D2_KC>
GB_D void d2_compute(const double* __restrict__ stage, int wr, int wc, int lr, int lc,
                     const double (&sl)[LB], D2Acc (&acc)[LB]) {
    constexpr int D2_AS = D2_KC + 4;
    const double* bst = stage + LB * 2 * D2_TM * D2_AS;
    constexpr int NA = MU ? LB : 1;
#pragma unroll
    for (int kk = 0; kk < D2_KC / 4; ++kk) {
        double are[NA], aim[NA], bre[LB], bim[LB];
#pragma unroll
        for (int q = 0; q < NA; ++q) {
            are[q] = stage[((q * 2 + 0) * D2_TM + wr * 8 + lr) * D2_AS + kk * 4 + lc];
            aim[q] = stage[((q * 2 + 1) * D2_TM + wr * 8 + lr) * D2_AS + kk * 4 + lc];
        }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            bre[q] = bst[((q * 2 + 0) * D2_KC + kk * 4 + lc) * D2_BS + wc * 8 + lr];
            bim[q] = bst[((q * 2 + 1) * D2_KC + kk * 4 + lc) * D2_BS + wc * 8 + lr];
        }
        const double nai = -aim[0];
        // first and second product of every accumulator are issued 2*LB instructions apart
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            dmma884(acc[l].re, are[0], bre[l]);
            dmma884(acc[l].im, are[0], bim[l]);
        }
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            dmma884(acc[l].re, nai, bim[l]);
            dmma884(acc[l].im, aim[0], bre[l]);
        }
        if (MU) {
            double bsr[LB], bsi[LB];
#pragma unroll
            for (int l = 0; l < LB - 1; ++l) {
                bsr[l] = sl[l] * bre[LB - 1];
                bsi[l] = sl[l] * bim[LB - 1];
            }
#pragma unroll
            for (int l = 0; l < LB - 1; ++l) {
                dmma884(acc[l].re, are[MU ? 1 + l : 0], bsr[l]);
                dmma884(acc[l].im, are[MU ? 1 + l : 0], bsi[l]);
            }
#pragma unroll
            for (int l = 0; l < LB - 1; ++l) {
                dmma884(acc[l].re, -aim[MU ? 1 + l : 0], bsi[l]);
                dmma884(acc[l].im, aim[MU ? 1 + l : 0], bsr[l]);
            }
        }
    }
}

// acc_l = sum_k A_0[r][k] B_l[k][c]  (+ s_l A_{1+l}[r][k] B_{LB-1}[k][c] for l < LB-1 if MU)
template <int LB, bool MU, int D2_KC>
GB_D void d2_tile_gemm(double* __restrict__ sm, const double* const (&A)[LB], size_t aplane, int lda,
                       const double* const (&B)[LB], size_t bplane, int ldb, int Nk,
                       const double (&sl)[LB], D2Acc (&acc)[LB]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3, wr = w >> 1, wc = w & 1;
    constexpr size_t SD = d2_stage_doubles(LB, D2_KC);
#pragma unroll
    for (int l = 0; l < LB; ++l) acc[l].re[0] = acc[l].re[1] = acc[l].im[0] = acc[l].im[1] = 0.0;
#pragma unroll
    for (int s = 0; s < D2_ST - 1; ++s) {
        if (s < Nk) d2_issue<LB, D2_KC>(sm + s * SD, A, aplane, lda, B, bplane, ldb, s * D2_KC);
        cp_async_commit();
    }
    for (int c = 0; c < Nk; ++c) {
        cp_async_wait<D2_ST - 2>();
        __syncthreads();
        const int nx = c + D2_ST - 1;
        if (nx < Nk) d2_issue<LB, D2_KC>(sm + (nx % D2_ST) * SD, A, aplane, lda, B, bplane, ldb, nx * D2_KC);
        cp_async_commit();
        d2_compute<LB, MU, D2_KC>(sm + (c % D2_ST) * SD, wr, wc, lr, lc, sl, acc);
    }
    __syncthreads();
}

// Single-operand tile GEMM of the chain kernels with 4-way split of the contraction index over the warps.
// ncu on dense2_chain (profiles/r1_s5_ncu_c5_chain_raw.csv): DMMA pipe 51 % active, top stall "wait" --
// with one 8x8 output tile per warp there are only two dependent accumulator chains (re, im) per warp and
// the latency of a dependent DMMA (~100 clk) caps the rate.  Here a warp owns a 16x16 block (2x2 DMMA
// tiles = 8 independent chains, operand fragments reused twice) over one quarter of every staged chunk;
// the four partial sums meet in shared memory once per GEMM.  Result in the same per-thread mapping as
// d2_tile_gemm (row wr*8+lr, columns wc*8+2lc+{0,1}).
GB_D void d2_tile_gemm_k4(double* __restrict__ sm, const double* __restrict__ A, size_t aplane, int lda,
                          const double* __restrict__ B, size_t bplane, int ldb, int Nk, D2Acc& out) {
    constexpr int KC = D2_KC_F, AS = KC + 4, RS = 24;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    const int kq = w & 3, rh = w >> 2;
    constexpr size_t SD = d2_stage_doubles(1, KC);
    const double* const Aa[1] = {A};
    const double* const Ba[1] = {B};
    D2Acc acc[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j].re[0] = acc[i][j].re[1] = acc[i][j].im[0] = acc[i][j].im[1] = 0.0;
#pragma unroll
    for (int s = 0; s < D2_ST - 1; ++s) {
        if (s < Nk) d2_issue<1, KC>(sm + s * SD, Aa, aplane, lda, Ba, bplane, ldb, s * KC);
        cp_async_commit();
    }
    for (int c = 0; c < Nk; ++c) {
        cp_async_wait<D2_ST - 2>();
        __syncthreads();
        const int nx = c + D2_ST - 1;
        if (nx < Nk) d2_issue<1, KC>(sm + (nx % D2_ST) * SD, Aa, aplane, lda, Ba, bplane, ldb, nx * KC);
        cp_async_commit();
        const double* st = sm + (c % D2_ST) * SD;
        const double* bst = st + 2 * D2_TM * AS;
#pragma unroll
        for (int kk = 0; kk < KC / 16; ++kk) {
            const int ko = kq * (KC / 4) + kk * 4 + lc;
            double are[2], aim[2], bre[2], bim[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                are[i] = st[(0 * D2_TM + rh * 16 + i * 8 + lr) * AS + ko];
                aim[i] = st[(1 * D2_TM + rh * 16 + i * 8 + lr) * AS + ko];
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                bre[j] = bst[(0 * KC + ko) * D2_BS + j * 8 + lr];
                bim[j] = bst[(1 * KC + ko) * D2_BS + j * 8 + lr];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    dmma884(acc[i][j].re, are[i], bre[j]);
                    dmma884(acc[i][j].im, are[i], bim[j]);
                }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double nai = -aim[i];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    dmma884(acc[i][j].re, nai, bim[j]);
                    dmma884(acc[i][j].im, aim[i], bre[j]);
                }
            }
        }
    }
    __syncthreads();   // every warp is done with the stage buffers: reuse them for the partial sums
    // red[kq][plane][row 0..31][RS]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = rh * 16 + i * 8 + lr, col = j * 8 + 2 * lc;
            *reinterpret_cast<double2*>(&sm[((kq * 2 + 0) * D2_TM + row) * RS + col]) = make_double2(acc[i][j].re[0], acc[i][j].re[1]);
            *reinterpret_cast<double2*>(&sm[((kq * 2 + 1) * D2_TM + row) * RS + col]) = make_double2(acc[i][j].im[0], acc[i][j].im[1]);
        }
    __syncthreads();
    {
        const int wr = w >> 1, wc = w & 1;
        const int row = wr * 8 + lr, col = wc * 8 + 2 * lc;
        double2 r = make_double2(0.0, 0.0), im = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double2 a = *reinterpret_cast<const double2*>(&sm[((q * 2 + 0) * D2_TM + row) * RS + col]);
            const double2 b = *reinterpret_cast<const double2*>(&sm[((q * 2 + 1) * D2_TM + row) * RS + col]);
            r.x += a.x; r.y += a.y; im.x += b.x; im.y += b.y;
        }
        out.re[0] = r.x; out.re[1] = r.y; out.im[0] = im.x; out.im[1] = im.y;
    }
    __syncthreads();
}

// H_n = M_0 + sum_l a_l M_{1+l} for all matrix elements, distributed over the grid
GB_D void d2_form(const DevP& p, const double* __restrict__ Mall, double* __restrict__ out, size_t hplane, int n) {
    // ncu (profiles/r1_s5_ncu_c5_chain_*): written as one element per loop trip this took 18 % of the chain
    // kernel on load latency alone; now 16-byte loads, 8 of them in flight per thread and operator.
    constexpr int U = 8;
    const size_t tot2 = hplane;   // double2 elements of one operator (two planes of hplane doubles)
    const double2* __restrict__ M2 = reinterpret_cast<const double2*>(Mall);
    double2* __restrict__ O2 = reinterpret_cast<double2*>(out);
    const size_t stride = (size_t)gridDim.x * D2_THREADS;
    for (size_t e0 = (size_t)blockIdx.x * D2_THREADS + threadIdx.x; e0 < tot2; e0 += U * stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t e = e0 + u * stride;
            v[u] = e < tot2 ? __ldg(&M2[e]) : make_double2(0.0, 0.0);
        }
        for (int l = 0; l < p.L; ++l) {
            double al = p.eps[l * p.NT + n];
            if (p.shape) al *= p.shape[l * p.NT + n];
            double2 t[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t e = e0 + u * stride;
                t[u] = e < tot2 ? __ldg(&M2[(size_t)(1 + l) * tot2 + e]) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u].x = fma(al, t[u].x, v[u].x);
                v[u].y = fma(al, t[u].y, v[u].y);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t e = e0 + u * stride;
            if (e < tot2) O2[e] = v[u];
        }
    }
}

// ---------------------------------------------------------------------------
// chain kernel: forward sweep (BWD = false; reference src/optimize.jl:720-751) or the chi chain of the
// Krylov-form backward (BWD = true; see dense_chain in dense.cuh)
// ---------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(D2_THREADS, 1) dense2_chain(DevP p, DenseDev d, Dense2Dev d2, KryDev kd) {
    if (BWD && !(*kd.ok)) return;   // uniform over the grid
    cgx::grid_group grid = cgx::this_grid();
    extern __shared__ __align__(16) double dsm[];
    __shared__ double s_gb[4][D2_TN];
    const int Np = d.Np, Kp = d.Kp, NT = p.NT;
    const size_t splane = (size_t)Np * Kp, hplane = (size_t)Np * Np;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3, wr = w >> 1, wc = w & 1;
    const int NkF = Np / D2_KC_F;
    const bool gb = BWD ? (p.gb_kind != 0 && p.lambda_b != 0.0) : (p.gb_kind != 0);
    double* cur = BWD ? kd.kcur : d.cur;
    double* nxt = BWD ? kd.kcur2 : d2.cur2;
    double* const cur_first = cur;
    const double* Hall = BWD ? d.Ha : d.Hf;
    double* terms = BWD ? kd.BT : kd.FT;

    // J_b contribution of the state in `cur` with weight wgt: Re <psi|D|psi> per trajectory
    auto gb_point = [&](double wgt) {
        for (int t = blockIdx.x; t < d2.ntiles; t += gridDim.x) {
            const int rt = t / d2.tilesC, ct = t % d2.tilesC;
            const int r0 = rt * D2_TM, c0 = ct * D2_TN;
            D2Acc acc[1];
            d2_tile_gemm_k4(dsm, d.Dm + (size_t)r0 * Np, hplane, Np, cur + c0, splane, Kp, NkF, acc[0]);
            const int row = r0 + wr * 8 + lr;
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const size_t off = (size_t)row * Kp + c0 + wc * 8 + 2 * lc + e;
                v[e] = cur[off] * acc[0].re[e] + cur[splane + off] * acc[0].im[e];
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                v[e] += __shfl_xor_sync(0xffffffffu, v[e], 4);
                v[e] += __shfl_xor_sync(0xffffffffu, v[e], 8);
                v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
            }
            if (lr == 0) { s_gb[wr][wc * 8 + 2 * lc] = v[0]; s_gb[wr][wc * 8 + 2 * lc + 1] = v[1]; }
            __syncthreads();
            if (threadIdx.x < D2_TN) {
                const double sacc = s_gb[0][threadIdx.x] + s_gb[1][threadIdx.x] + s_gb[2][threadIdx.x] + s_gb[3][threadIdx.x];
                d.jbpart[(size_t)rt * Kp + c0 + threadIdx.x] += wgt * sacc;
            }
            __syncthreads();
        }
    };

    const int nfirst = BWD ? NT - 1 : 0;
    d2_form(p, Hall, d2.Hn[nfirst & 1], hplane, nfirst);
    if (BWD) {   // slot 0 of the last step = chi(T)
        double* s0 = terms + (size_t)(NT - 1) * kd.MT * 2 * splane;
        for (size_t e = (size_t)blockIdx.x * D2_THREADS + threadIdx.x; e < 2 * splane; e += (size_t)gridDim.x * D2_THREADS)
            s0[e] = cur[e];
    }
    grid.sync();
    for (int it = 0; it < NT; ++it) {
        const int n = BWD ? NT - 1 - it : it;
        const int nnext = BWD ? n - 1 : n + 1;
        const bool has_next = BWD ? n > 0 : n + 1 < NT;
        const double dt = p.tlist[n + 1] - p.tlist[n];
        const double* Hn = d2.Hn[n & 1];
        if (!BWD && gb) gb_point(n == 0 ? 0.5 * (p.tlist[1] - p.tlist[0]) : 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]));
        int m, s;
        const double* gw;   // weights of the Taylor terms in the new state (dense.cuh: economised polynomial)
        dense_plan(p, d, n, dt, m, s, gw, d.econ && kd.on);
        if (!BWD && p.grad_method != 0 && p.taylor_check && m > p.taylor_max_order && blockIdx.x == 0 && threadIdx.x == 0)
            p.flags->taylor_fail = 1;
        const bool kry = kd.on && s == 0 && m <= kd.MT;
        double* slots = kry ? terms + (size_t)n * kd.MT * 2 * splane : nullptr;
        const int nsub = 1 << s;
        const double dts = dt / nsub;
        for (int sub = 0; sub < nsub; ++sub) {
            const bool last = sub == nsub - 1;
            for (int j = 1; j <= m; ++j) {
                const double* src = j == 1 ? cur : (kry ? slots + (size_t)(j - 1) * 2 * splane : ((j - 1) & 1 ? d.T1 : d.T0));
                double* dst = kry ? slots + (size_t)j * 2 * splane : ((j & 1) ? d.T1 : d.T0);
                const double x = dts / j;
                const bool fin = j == m && last;
                if (fin && has_next) d2_form(p, Hall, d2.Hn[nnext & 1], hplane, nnext);
                for (int t = blockIdx.x; t < d2.ntiles; t += gridDim.x) {
                    const int r0 = (t / d2.tilesC) * D2_TM, c0 = (t % d2.tilesC) * D2_TN;
                    const int row = r0 + wr * 8 + lr;
                    const int kc = c0 + wc * 8 + 2 * lc;
                    const size_t off = (size_t)row * Kp + kc;
                    const double* base = j == 1 ? cur : nxt;
                    // running sum of this thread's elements: loaded before the GEMM so that the latency is hidden
                    const double2 b_r = *reinterpret_cast<const double2*>(&base[off]);
                    const double2 b_i = *reinterpret_cast<const double2*>(&base[splane + off]);
                    D2Acc acc[1];
                    d2_tile_gemm_k4(dsm, Hn + (size_t)r0 * Np, hplane, Np, src + c0, splane, Kp, NkF, acc[0]);
                    double tr[2], ti[2], vr[2], vi[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {   // forward: t = (-i x) * res ; backward: t = (+i x) * res
                        tr[e] = BWD ? -x * acc[0].im[e] : x * acc[0].im[e];
                        ti[e] = BWD ? x * acc[0].re[e] : -x * acc[0].re[e];
                    }
                    vr[0] = fma(gw[j], tr[0], b_r.x); vr[1] = fma(gw[j], tr[1], b_r.y);
                    vi[0] = fma(gw[j], ti[0], b_i.x); vi[1] = fma(gw[j], ti[1], b_i.y);
                    if (j < m) {
                        *reinterpret_cast<double2*>(&dst[off]) = make_double2(tr[0], tr[1]);
                        *reinterpret_cast<double2*>(&dst[splane + off]) = make_double2(ti[0], ti[1]);
                    }
                    if (BWD && fin && gb && n > 0) {
                        // chi += lambda_b * 0.5 (t_{n+1} - t_{n-1}) / rho * xi(Psi(t_{n-1})), xi = -D Psi  (optimize.jl:897-908)
                        const double f = p.lambda_b * 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]);
                        const double* st = d.store + (size_t)n * 2 * splane;
                        D2Acc a1[1];
                        d2_tile_gemm_k4(dsm, d.Dm + (size_t)r0 * Np, hplane, Np, st + c0, splane, Kp, NkF, a1[0]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = kc + e;
                            if (k < p.K) {
                                const double fk = f / p.rho[k];
                                vr[e] -= fk * a1[0].re[e];
                                vi[e] -= fk * a1[0].im[e];
                            }
                        }
                    }
                    *reinterpret_cast<double2*>(&nxt[off]) = make_double2(vr[0], vr[1]);
                    *reinterpret_cast<double2*>(&nxt[splane + off]) = make_double2(vi[0], vi[1]);
                    if (fin) {
                        if (!BWD) {
                            double* st = d.store + (size_t)(n + 1) * 2 * splane;
                            __stcs(reinterpret_cast<double2*>(&st[off]), make_double2(vr[0], vr[1]));
                            __stcs(reinterpret_cast<double2*>(&st[splane + off]), make_double2(vi[0], vi[1]));
                        } else if (n > 0) {   // slot 0 of the next (earlier) step = chi(t_{n-1})
                            double* s0 = terms + (size_t)(n - 1) * kd.MT * 2 * splane;
                            *reinterpret_cast<double2*>(&s0[off]) = make_double2(vr[0], vr[1]);
                            *reinterpret_cast<double2*>(&s0[splane + off]) = make_double2(vi[0], vi[1]);
                        }
                    }
                }
                grid.sync();
            }
            double* tmp = cur; cur = nxt; nxt = tmp;
        }
    }
    if (!BWD) {
        if (gb) gb_point(0.5 * (p.tlist[NT] - p.tlist[NT - 1]));
        // final state must be in d.cur (read by dense_tau / dense_boundary / read-backs)
        if (cur != cur_first) {
            for (size_t e = (size_t)blockIdx.x * D2_THREADS + threadIdx.x; e < 2 * splane; e += (size_t)gridDim.x * D2_THREADS)
                cur_first[e] = cur[e];
        }
    }
}

// ---------------------------------------------------------------------------
// backward sweep fused with the gradient contraction (reference src/optimize.jl:880-911)
// ---------------------------------------------------------------------------
template <int LB>
__global__ void __launch_bounds__(D2_THREADS, 1) dense2_backward(DevP p, DenseDev d, Dense2Dev d2) {
    constexpr int L = LB - 1;
    if (d.kry_ok && *d.kry_ok) return;   // the Krylov-form kernels (dense_kry.cuh) serve this call; uniform over the grid
    cgx::grid_group grid = cgx::this_grid();
    extern __shared__ __align__(16) double dsm[];
    __shared__ double s_buf[32 * (L > 0 ? L : 1)];
    const int Np = d.Np, Kp = d.Kp, Cb = d.Cb, NT = p.NT;
    const size_t splane = (size_t)Np * Kp, bplane = (size_t)Np * Cb, hplane = (size_t)Np * Np;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3, wr = w >> 1, wc = w & 1;
    const int NkB = Np / D2_KC_B, NkF = Np / D2_KC_F;
    const bool gb = p.gb_kind != 0 && p.lambda_b != 0.0;
    double* cur = d.bcur;
    double* nxt = d2.nxt;

    d2_form(p, d.Ha, d2.Hn[(NT - 1) & 1], hplane, NT - 1);
    grid.sync();
    for (int n = NT - 1; n >= 0; --n) {
        const double dt = p.tlist[n + 1] - p.tlist[n];
        const double* Hn = d2.Hn[n & 1];
        double sl[LB];
#pragma unroll
        for (int l = 0; l < LB; ++l) sl[l] = (l < L && p.dshape) ? p.dshape[l * NT + n] : 1.0;
        int m, s;
        dense_plan(p, d, n, dt, m, s);
        const int nsub = 1 << s;
        const double dts = dt / nsub;
        for (int sub = 0; sub < nsub; ++sub) {
            const bool last = sub == nsub - 1;
            for (int j = 1; j <= m; ++j) {
                const double* src = j == 1 ? cur : ((j - 1) & 1 ? d.T1 : d.T0);
                double* dst = (j & 1) ? d.T1 : d.T0;
                const double x = dts / j;
                const bool fin = (j == m) && last;
                if (fin && n > 0) d2_form(p, d.Ha, d2.Hn[(n - 1) & 1], hplane, n - 1);
                double gv[L > 0 ? L : 1];
#pragma unroll
                for (int l = 0; l < L; ++l) gv[l] = 0.0;
                for (int t = blockIdx.x; t < d2.ntiles; t += gridDim.x) {
                    const int r0 = (t / d2.tilesC) * D2_TM, c0 = (t % d2.tilesC) * D2_TN;
                    const double* A[LB];
                    const double* B[LB];
                    A[0] = Hn + (size_t)r0 * Np;
#pragma unroll
                    for (int l = 0; l < L; ++l) A[1 + l] = d.Ha + (size_t)(1 + l) * 2 * hplane + (size_t)r0 * Np;
#pragma unroll
                    for (int l = 0; l < LB; ++l) B[l] = src + (size_t)l * Kp + c0;
                    const int row = r0 + wr * 8 + lr;
                    const int kc = c0 + wc * 8 + 2 * lc;             // trajectory index of element e = 0
                    const double* base = j == 1 ? cur : nxt;
                    double2 b_r[LB], b_i[LB];                        // running sums, loaded ahead of the GEMM
#pragma unroll
                    for (int l = 0; l < LB; ++l) {
                        const size_t off = (size_t)row * Cb + (size_t)l * Kp + kc;
                        b_r[l] = *reinterpret_cast<const double2*>(&base[off]);
                        b_i[l] = *reinterpret_cast<const double2*>(&base[bplane + off]);
                    }
                    D2Acc acc[LB];
                    d2_tile_gemm<LB, true, D2_KC_B>(dsm, A, hplane, Np, B, bplane, Cb, NkB, sl, acc);
                    double chi_r[2], chi_i[2];
#pragma unroll
                    for (int l = 0; l < LB; ++l) {
                        const size_t off = (size_t)row * Cb + (size_t)l * Kp + kc;
                        double tr[2], ti[2], vr[2], vi[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {   // t = (+i x) * res
                            tr[e] = -x * acc[l].im[e];
                            ti[e] = x * acc[l].re[e];
                        }
                        vr[0] = b_r[l].x + tr[0]; vr[1] = b_r[l].y + tr[1];
                        vi[0] = b_i[l].x + ti[0]; vi[1] = b_i[l].y + ti[1];
                        if (j < m) {
                            *reinterpret_cast<double2*>(&dst[off]) = make_double2(tr[0], tr[1]);
                            *reinterpret_cast<double2*>(&dst[bplane + off]) = make_double2(ti[0], ti[1]);
                        }
                        if (fin && l < L) {
                            // tau_grad[k][n,l] = rho_k <chi'_lk | Psi_k(t_{n-1})>  (optimize.jl:893-895); then resetgradvec! (:896)
                            const double* st = d.store + (size_t)n * 2 * splane;
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int k = kc + e;
                                if (k < p.K) {
                                    const size_t so = (size_t)row * Kp + k;
                                    gv[l < L ? l : 0] += p.rho[k] * (vr[e] * __ldg(&st[so]) + vi[e] * __ldg(&st[splane + so]));
                                }
                                vr[e] = 0.0; vi[e] = 0.0;
                            }
                        }
                        if (l == L) { chi_r[0] = vr[0]; chi_r[1] = vr[1]; chi_i[0] = vi[0]; chi_i[1] = vi[1]; }
                        else {
                            *reinterpret_cast<double2*>(&nxt[off]) = make_double2(vr[0], vr[1]);
                            *reinterpret_cast<double2*>(&nxt[bplane + off]) = make_double2(vi[0], vi[1]);
                        }
                    }
                    if (fin && gb && n > 0) {
                        // chi += lambda_b * 0.5 (t_{n+1} - t_{n-1}) / rho * xi(Psi(t_{n-1})), xi = -D Psi  (optimize.jl:897-908)
                        const double f = p.lambda_b * 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]);
                        const double* st = d.store + (size_t)n * 2 * splane;
                        const double* A1[1] = {d.Dm + (size_t)r0 * Np};
                        const double* B1[1] = {st + c0};
                        const double one[1] = {1.0};
                        D2Acc a1[1];
                        d2_tile_gemm<1, false, D2_KC_F>(dsm, A1, hplane, Np, B1, splane, Kp, NkF, one, a1);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = kc + e;
                            if (k < p.K) {
                                const double fk = f / p.rho[k];
                                chi_r[e] -= fk * a1[0].re[e];
                                chi_i[e] -= fk * a1[0].im[e];
                            }
                        }
                    }
                    {
                        const size_t off = (size_t)row * Cb + (size_t)L * Kp + kc;
                        *reinterpret_cast<double2*>(&nxt[off]) = make_double2(chi_r[0], chi_r[1]);
                        *reinterpret_cast<double2*>(&nxt[bplane + off]) = make_double2(chi_i[0], chi_i[1]);
                    }
                }
                if (fin) {
                    if (L > 0) {
                        block_sum<(L > 0 ? L : 1)>(gv, s_buf);
                        if (threadIdx.x == 0)
#pragma unroll
                            for (int l = 0; l < L; ++l) p.partial[(size_t)blockIdx.x * L * NT + (size_t)l * NT + n] = gv[l];
                    }
                }
                grid.sync();
            }
            double* tmp = cur; cur = nxt; nxt = tmp;
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct Dense2Plan {
    Dense2Dev d2;
    bool on;
    int grid;
    size_t smemF, smemB;
    size_t smemM;   // dense2_chain_multi<BWD, NS>
    Dense2Plan() : on(false), grid(0), smemF(0), smemB(0), smemM(0) { memset(&d2, 0, sizeof d2); }
};

template <int LB>
inline cudaError_t dense2_attr_b(size_t smem) {
    return cudaFuncSetAttribute(dense2_backward<LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

inline int dense2_setup(Dense2Plan& q, DensePlan& dp, DevP& p, std::vector<void*>& allocs, std::string& err) {
    DenseDev& d = dp.d;
    q.on = false;
    const int L = p.L, LB = L + 1;
    if (LB > D2_LBMAX) return 0;
    if (d.Kp % D2_TN != 0) {
        // pad the trajectory block to a multiple of the column tile: only when the caller did so (setup decides Kp)
        return 0;
    }
    Dense2Dev& d2 = q.d2;
    d2.tilesR = d.Np / D2_TM;
    d2.tilesC = d.Kp / D2_TN;
    d2.ntiles = d2.tilesR * d2.tilesC;
    const bool force = getenv("GRAPE_B200_DENSE2") && atoi(getenv("GRAPE_B200_DENSE2")) == 1;
    const bool off = getenv("GRAPE_B200_DENSE2") && atoi(getenv("GRAPE_B200_DENSE2")) == 0;
    // few tiles: the strip kernels (dense.cuh) use more SMs -- if they can run at all
    if ((off || (!force && d2.ntiles < 64)) && dp.strip_ok) return 0;
    q.smemF = sizeof(double) * d2_stage_doubles(1, D2_KC_F) * D2_ST;
    q.smemB = std::max(sizeof(double) * d2_stage_doubles(LB, D2_KC_B) * D2_ST, q.smemF);
    if (q.smemB > 200 * 1024) return 0;
    cudaError_t e = cudaFuncSetAttribute(dense2_chain<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.smemF);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dense2_chain<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.smemF);
    if (e == cudaSuccess) {
        switch (LB) {
            case 2: e = dense2_attr_b<2>(q.smemB); break;
            case 3: e = dense2_attr_b<3>(q.smemB); break;
            case 4: e = dense2_attr_b<4>(q.smemB); break;
            case 5: e = dense2_attr_b<5>(q.smemB); break;
            default: return 0;
        }
    }
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute failed (dense2): ") + cudaGetErrorString(e); return GRAPE_B200_ECUDA; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    q.grid = std::min(d2.ntiles, sms);
    const size_t hplane = (size_t)d.Np * d.Np, splane = (size_t)d.Np * d.Kp, bplane = (size_t)d.Np * d.Cb;
    auto ald = [&](double** dst, size_t n) -> int {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, n * sizeof(double)) != cudaSuccess) { err = "cudaMalloc failed (dense2)"; return GRAPE_B200_ECUDA; }
        allocs.push_back(ptr);
        cudaMemset(ptr, 0, n * sizeof(double));
        *dst = static_cast<double*>(ptr);
        return 0;
    };
    int rc;
    if ((rc = ald(&d2.Hn[0], 2 * hplane))) return rc;
    if ((rc = ald(&d2.Hn[1], 2 * hplane))) return rc;
    if ((rc = ald(&d2.nxt, 2 * bplane))) return rc;
    if ((rc = ald(&d2.cur2, 2 * splane))) return rc;
    // jbpart is indexed [row tile][Kp] here: make sure it is large enough
    if ((size_t)d2.tilesR > (size_t)dp.gridF) {
        if ((rc = ald(&d.jbpart, (size_t)d2.tilesR * d.Kp))) return rc;
    }
    if (q.grid > p.KB) {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, (size_t)q.grid * L * p.NT * sizeof(double)) != cudaSuccess) { err = "cudaMalloc failed (dense2)"; return GRAPE_B200_ECUDA; }
        allocs.push_back(ptr);
        p.partial = static_cast<double*>(ptr);
    }
    p.KB = q.grid;
    q.on = true;
    return 0;
}

inline void dense2_chain_multi_launch(Dense2Plan& q, bool bwd, int ns, void** args, cudaStream_t st);

inline void dense2_run_forward(Dense2Plan& q, DensePlan& dp, const DevP& p, cudaStream_t st, int64_t& launches) {
    DenseDev& d = dp.d;
    const size_t splane = (size_t)d.Np * d.Kp;
    cudaMemcpyAsync(d.cur, d.psi0, 2 * splane * sizeof(double), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(d.store, d.psi0, 2 * splane * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (p.gb_kind) cudaMemsetAsync(d.jbpart, 0, (size_t)q.d2.tilesR * d.Kp * sizeof(double), st);
    DevP pp = p;
    void* args[] = {&pp, &d, &q.d2, &dp.kd};
    if (d.nstrip > 1) {
        dense_run_preform(dp, p, false, st, launches);
        dense2_chain_multi_launch(q, false, d.nstrip, args, st);
    } else {
        cudaLaunchCooperativeKernel((void*)dense2_chain<false>, dim3(q.grid), dim3(D2_THREADS), args, q.smemF, st);
    }
    dense_tau<<<p.K, 256, 0, st>>>(p, d, q.d2.tilesR);
    launches += 2;
}

inline void dense2_run_backward(Dense2Plan& q, DensePlan& dp, const DevP& p, const cplx* chi_host, cudaStream_t st, int64_t& launches) {
    DenseDev& d = dp.d;
    const size_t bplane = (size_t)d.Np * d.Cb;
    cudaMemsetAsync(d.bcur, 0, 2 * bplane * sizeof(double), st);
    dense_boundary<<<p.K, 256, 0, st>>>(p, d, chi_host, dp.kd.on ? dp.kd.kcur : nullptr, nullptr);
    DevP pp = p;
    void* args[] = {&pp, &d, &q.d2};
    void* fn = nullptr;
    switch (p.L + 1) {
        case 2: fn = (void*)dense2_backward<2>; break;
        case 3: fn = (void*)dense2_backward<3>; break;
        case 4: fn = (void*)dense2_backward<4>; break;
        default: fn = (void*)dense2_backward<5>; break;
    }
    cudaLaunchCooperativeKernel(fn, dim3(q.grid), dim3(D2_THREADS), args, q.smemB, st);
    launches += 2;
    if (dp.kd.on) {
        void* cargs[] = {&pp, &d, &q.d2, &dp.kd};
        if (d.nstrip > 1) {
            if (d.preA != d.preF) dense_run_preform(dp, p, true, st, launches);
            dense2_chain_multi_launch(q, true, d.nstrip, cargs, st);
        } else {
            cudaLaunchCooperativeKernel((void*)dense2_chain<true>, dim3(q.grid), dim3(D2_THREADS), cargs, q.smemF, st);
        }
        launches += 1;
    }
}

// ---------------------------------------------------------------------------
// Tiled chain with NS = 2 or 3 Taylor terms per grid barrier (see dense_chain<BWD, NS> in dense.cuh):
//   T_{j+q} = (-+i dt)^q / ((j+1)..(j+q)) H_n^q T_j,  q = 1..NS,
// one staged chunk of the operand block T_j feeds NS products with the tiles of H_n, H_n^2 [, H_n^3], which are read
// from the generators pre-formed for all steps of the call (dense_preform; no per-step formation in this kernel).
// ---------------------------------------------------------------------------
template <int NS>
__host__ __device__ constexpr size_t d2m_stage_doubles() { return (size_t)NS * 2 * D2_TM * (D2_KC_F + 4) + 2 * D2_KC_F * D2_BS; }

// stage layout: A_q plane pl at ((q*2+pl)*TM + r)*AS, then B plane pl at ((pl)*KC + kr)*BS
template <int NS>
GB_D void d2m_issue(double* __restrict__ stage, const double* __restrict__ A, size_t aplane, int lda,
                    const double* __restrict__ B, size_t bplane, int ldb, int k0, int nt) {
    constexpr int KC = D2_KC_F, AS = KC + 4, HS = KC / 2, ASEG = 2 * D2_TM * HS, TS = D2_TN / 2, BSEG = 2 * KC * TS;
    double* bst = stage + NS * 2 * D2_TM * AS;
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        if (q < nt) {
#pragma unroll
            for (int e0 = 0; e0 < ASEG; e0 += D2_THREADS) {
                const int e = e0 + threadIdx.x;
                const int s = e % HS, r = (e / HS) % D2_TM, pl = e / (HS * D2_TM);
                if (ASEG % D2_THREADS == 0 || e < ASEG)
                    cp_async16(stage + ((q * 2 + pl) * D2_TM + r) * AS + 2 * s, A + (size_t)(2 * q + pl) * aplane + (size_t)r * lda + k0 + 2 * s);
            }
        }
    }
#pragma unroll
    for (int e0 = 0; e0 < BSEG; e0 += D2_THREADS) {
        const int e = e0 + threadIdx.x;
        const int s = e % TS, kr = (e / TS) % KC, pl = e / (TS * KC);
        if (BSEG % D2_THREADS == 0 || e < BSEG)
            cp_async16(bst + (pl * KC + kr) * D2_BS + 2 * s, B + pl * bplane + (size_t)(k0 + kr) * ldb + 2 * s);
    }
}

// out[q] = sum_k A_q[r][k] B[k][c], q < nt; warp = 16x16 block over one quarter of every staged chunk (as
// d2_tile_gemm_k4); result in the per-thread mapping row wr*8+lr, columns wc*8+2lc+{0,1}.
template <int NS>
GB_D void d2m_tile_gemm(double* __restrict__ sm, const double* __restrict__ A, size_t aplane, int lda,
                        const double* __restrict__ B, size_t bplane, int ldb, int Nk, int nt, D2Acc (&out)[NS]) {
    constexpr int KC = D2_KC_F, AS = KC + 4, RS = 24;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    const int kq = w & 3, rh = w >> 2;
    constexpr size_t SD = d2m_stage_doubles<NS>();
    D2Acc acc[NS][2][2];
#pragma unroll
    for (int q = 0; q < NS; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[q][i][j].re[0] = acc[q][i][j].re[1] = acc[q][i][j].im[0] = acc[q][i][j].im[1] = 0.0;
#pragma unroll
    for (int s = 0; s < D2_ST - 1; ++s) {
        if (s < Nk) d2m_issue<NS>(sm + s * SD, A, aplane, lda, B, bplane, ldb, s * KC, nt);
        cp_async_commit();
    }
    for (int c = 0; c < Nk; ++c) {
        cp_async_wait<D2_ST - 2>();
        __syncthreads();
        const int nx = c + D2_ST - 1;
        if (nx < Nk) d2m_issue<NS>(sm + (nx % D2_ST) * SD, A, aplane, lda, B, bplane, ldb, nx * KC, nt);
        cp_async_commit();
        const double* st = sm + (c % D2_ST) * SD;
        const double* bst = st + NS * 2 * D2_TM * AS;
#pragma unroll
        for (int kk = 0; kk < KC / 16; ++kk) {
            const int ko = kq * (KC / 4) + kk * 4 + lc;
            double bre[2], bim[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                bre[j] = bst[(0 * KC + ko) * D2_BS + j * 8 + lr];
                bim[j] = bst[(1 * KC + ko) * D2_BS + j * 8 + lr];
            }
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                if (q < nt) {
                    double are[2], aim[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        are[i] = st[((q * 2 + 0) * D2_TM + rh * 16 + i * 8 + lr) * AS + ko];
                        aim[i] = st[((q * 2 + 1) * D2_TM + rh * 16 + i * 8 + lr) * AS + ko];
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            dmma884(acc[q][i][j].re, are[i], bre[j]);
                            dmma884(acc[q][i][j].im, are[i], bim[j]);
                        }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double nai = -aim[i];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            dmma884(acc[q][i][j].re, nai, bim[j]);
                            dmma884(acc[q][i][j].im, aim[i], bre[j]);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();   // every warp is done with the stage buffers: reuse them for the partial sums
    // red[q][kq][plane][row 0..31][RS]
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        if (q < nt) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int row = rh * 16 + i * 8 + lr, col = j * 8 + 2 * lc;
                    *reinterpret_cast<double2*>(&sm[(((q * 4 + kq) * 2 + 0) * D2_TM + row) * RS + col]) = make_double2(acc[q][i][j].re[0], acc[q][i][j].re[1]);
                    *reinterpret_cast<double2*>(&sm[(((q * 4 + kq) * 2 + 1) * D2_TM + row) * RS + col]) = make_double2(acc[q][i][j].im[0], acc[q][i][j].im[1]);
                }
        }
    }
    __syncthreads();
    {
        const int wr = w >> 1, wc = w & 1;
        const int row = wr * 8 + lr, col = wc * 8 + 2 * lc;
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            double2 r = make_double2(0.0, 0.0), im = make_double2(0.0, 0.0);
            if (q < nt) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double2 a = *reinterpret_cast<const double2*>(&sm[(((q * 4 + t) * 2 + 0) * D2_TM + row) * RS + col]);
                    const double2 b = *reinterpret_cast<const double2*>(&sm[(((q * 4 + t) * 2 + 1) * D2_TM + row) * RS + col]);
                    r.x += a.x; r.y += a.y; im.x += b.x; im.y += b.y;
                }
            }
            out[q].re[0] = r.x; out[q].re[1] = r.y; out[q].im[0] = im.x; out[q].im[1] = im.y;
        }
    }
    __syncthreads();
}

template <bool BWD, int NS>
__global__ void __launch_bounds__(D2_THREADS, 1) dense2_chain_multi(DevP p, DenseDev d, Dense2Dev d2, KryDev kd) {
    if (BWD && !(*kd.ok)) return;   // uniform over the grid
    cgx::grid_group grid = cgx::this_grid();
    extern __shared__ __align__(16) double dsm[];
    __shared__ double s_gb[4][D2_TN];
    const int Np = d.Np, Kp = d.Kp, NT = p.NT;
    const size_t splane = (size_t)Np * Kp, hplane = (size_t)Np * Np;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3, wr = w >> 1, wc = w & 1;
    const int NkF = Np / D2_KC_F;
    const bool gb = BWD ? (p.gb_kind != 0 && p.lambda_b != 0.0) : (p.gb_kind != 0);
    double* cur = BWD ? kd.kcur : d.cur;
    double* nxt = BWD ? kd.kcur2 : d2.cur2;
    double* const cur_first = cur;
    const double* pre = BWD ? d.preA : d.preF;
    double* terms = BWD ? kd.BT : kd.FT;

    auto gb_point = [&](double wgt) {   // J_b contribution of the state in `cur`: Re <psi|D|psi> per trajectory
        for (int t = blockIdx.x; t < d2.ntiles; t += gridDim.x) {
            const int rt = t / d2.tilesC, ct = t % d2.tilesC;
            const int r0 = rt * D2_TM, c0 = ct * D2_TN;
            D2Acc acc[1];
            d2_tile_gemm_k4(dsm, d.Dm + (size_t)r0 * Np, hplane, Np, cur + c0, splane, Kp, NkF, acc[0]);
            const int row = r0 + wr * 8 + lr;
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const size_t off = (size_t)row * Kp + c0 + wc * 8 + 2 * lc + e;
                v[e] = cur[off] * acc[0].re[e] + cur[splane + off] * acc[0].im[e];
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                v[e] += __shfl_xor_sync(0xffffffffu, v[e], 4);
                v[e] += __shfl_xor_sync(0xffffffffu, v[e], 8);
                v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
            }
            if (lr == 0) { s_gb[wr][wc * 8 + 2 * lc] = v[0]; s_gb[wr][wc * 8 + 2 * lc + 1] = v[1]; }
            __syncthreads();
            if (threadIdx.x < D2_TN) {
                const double sacc = s_gb[0][threadIdx.x] + s_gb[1][threadIdx.x] + s_gb[2][threadIdx.x] + s_gb[3][threadIdx.x];
                d.jbpart[(size_t)rt * Kp + c0 + threadIdx.x] += wgt * sacc;
            }
            __syncthreads();
        }
    };

    if (BWD) {   // slot 0 of the last step = chi(T)
        double* s0 = terms + (size_t)(NT - 1) * kd.MT * 2 * splane;
        for (size_t e = (size_t)blockIdx.x * D2_THREADS + threadIdx.x; e < 2 * splane; e += (size_t)gridDim.x * D2_THREADS)
            s0[e] = cur[e];
        grid.sync();
    }
    for (int it = 0; it < NT; ++it) {
        const int n = BWD ? NT - 1 - it : it;
        const double dt = p.tlist[n + 1] - p.tlist[n];
        const double* Hq = pre + (size_t)n * (2 * NS) * hplane;   // planes {Re H, Im H, Re H^2, Im H^2, ..}
        if (!BWD && gb) gb_point(n == 0 ? 0.5 * (p.tlist[1] - p.tlist[0]) : 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]));
        int m, s;
        const double* gw;   // weights of the Taylor terms in the new state (dense.cuh: economised polynomial)
        dense_plan(p, d, n, dt, m, s, gw, d.econ && kd.on);
        if (!BWD && p.grad_method != 0 && p.taylor_check && m > p.taylor_max_order && blockIdx.x == 0 && threadIdx.x == 0)
            p.flags->taylor_fail = 1;
        const bool kry = kd.on && s == 0 && m <= kd.MT;
        double* slots = kry ? terms + (size_t)n * kd.MT * 2 * splane : nullptr;
        const bool multi = kry && m >= 2;
        const int nsub = 1 << s;
        const double dts = dt / nsub;
        for (int sub = 0; sub < nsub; ++sub) {
            const bool last = sub == nsub - 1;
            // stage = NS terms (multi) or one term; jt = index of the first term of the stage (1-based)
            for (int j = 1; j <= m; j += (multi ? NS : 1)) {
                const int nt = multi ? min(NS, m - j + 1) : 1;
                const double* src = j == 1 ? cur : (kry ? slots + (size_t)(j - 1) * 2 * splane : ((j - 1) & 1 ? d.T1 : d.T0));
                const bool fin = (j + nt - 1 == m) && last;
                double xq[NS];
                {
                    double x = 1.0;
#pragma unroll
                    for (int q = 0; q < NS; ++q) { x *= dts / (j + q); xq[q] = x; }
                }
                for (int t = blockIdx.x; t < d2.ntiles; t += gridDim.x) {
                    const int r0 = (t / d2.tilesC) * D2_TM, c0 = (t % d2.tilesC) * D2_TN;
                    const int row = r0 + wr * 8 + lr;
                    const int kc = c0 + wc * 8 + 2 * lc;
                    const size_t off = (size_t)row * Kp + kc;
                    const double* base = j == 1 ? cur : nxt;
                    const double2 b_r = *reinterpret_cast<const double2*>(&base[off]);
                    const double2 b_i = *reinterpret_cast<const double2*>(&base[splane + off]);
                    D2Acc acc[NS];
                    d2m_tile_gemm<NS>(dsm, Hq + (size_t)r0 * Np, hplane, Np, src + c0, splane, Kp, NkF, nt, acc);
                    double vr[2] = {b_r.x, b_r.y}, vi[2] = {b_i.x, b_i.y};
#pragma unroll
                    for (int q = 0; q < NS; ++q) {
                        if (q < nt) {
                            // f_q res: q = 0: (-+i) x res; q = 1: -x res; q = 2: (+-i) x res   (forward: -i, backward: +i)
                            double tr[2], ti[2];
                            const double x = xq[q];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                if (q == 0) { tr[e] = BWD ? -x * acc[q].im[e] : x * acc[q].im[e]; ti[e] = BWD ? x * acc[q].re[e] : -x * acc[q].re[e]; }
                                else if (q == 1) { tr[e] = -x * acc[q].re[e]; ti[e] = -x * acc[q].im[e]; }
                                else { tr[e] = BWD ? x * acc[q].im[e] : -x * acc[q].im[e]; ti[e] = BWD ? -x * acc[q].re[e] : x * acc[q].re[e]; }
                                vr[e] = fma(gw[j + q], tr[e], vr[e]);
                                vi[e] = fma(gw[j + q], ti[e], vi[e]);
                            }
                            if (j + q < m) {
                                double* dst = kry ? slots + (size_t)(j + q) * 2 * splane : (((j + q) & 1) ? d.T1 : d.T0);
                                *reinterpret_cast<double2*>(&dst[off]) = make_double2(tr[0], tr[1]);
                                *reinterpret_cast<double2*>(&dst[splane + off]) = make_double2(ti[0], ti[1]);
                            }
                        }
                    }
                    if (BWD && fin && gb && n > 0) {
                        // chi += lambda_b * 0.5 (t_{n+1} - t_{n-1}) / rho * xi(Psi(t_{n-1})), xi = -D Psi  (optimize.jl:897-908)
                        const double f = p.lambda_b * 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]);
                        const double* st = d.store + (size_t)n * 2 * splane;
                        D2Acc a1[1];
                        d2_tile_gemm_k4(dsm, d.Dm + (size_t)r0 * Np, hplane, Np, st + c0, splane, Kp, NkF, a1[0]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = kc + e;
                            if (k < p.K) {
                                const double fk = f / p.rho[k];
                                vr[e] -= fk * a1[0].re[e];
                                vi[e] -= fk * a1[0].im[e];
                            }
                        }
                    }
                    *reinterpret_cast<double2*>(&nxt[off]) = make_double2(vr[0], vr[1]);
                    *reinterpret_cast<double2*>(&nxt[splane + off]) = make_double2(vi[0], vi[1]);
                    if (fin) {
                        if (!BWD) {
                            double* st = d.store + (size_t)(n + 1) * 2 * splane;
                            __stcs(reinterpret_cast<double2*>(&st[off]), make_double2(vr[0], vr[1]));
                            __stcs(reinterpret_cast<double2*>(&st[splane + off]), make_double2(vi[0], vi[1]));
                        } else if (n > 0) {   // slot 0 of the next (earlier) step = chi(t_{n-1})
                            double* s0 = terms + (size_t)(n - 1) * kd.MT * 2 * splane;
                            *reinterpret_cast<double2*>(&s0[off]) = make_double2(vr[0], vr[1]);
                            *reinterpret_cast<double2*>(&s0[splane + off]) = make_double2(vi[0], vi[1]);
                        }
                    }
                }
                grid.sync();
            }
            double* tmp = cur; cur = nxt; nxt = tmp;
        }
    }
    if (!BWD) {
        if (gb) gb_point(0.5 * (p.tlist[NT] - p.tlist[NT - 1]));
        if (cur != cur_first) {   // final state must be in d.cur (read by dense_tau / dense_boundary / read-backs)
            for (size_t e = (size_t)blockIdx.x * D2_THREADS + threadIdx.x; e < 2 * splane; e += (size_t)gridDim.x * D2_THREADS)
                cur_first[e] = cur[e];
        }
    }
}

inline void dense2_chain_multi_launch(Dense2Plan& q, bool bwd, int ns, void** args, cudaStream_t st) {
    void* fn = bwd ? (ns == 3 ? (void*)dense2_chain_multi<true, 3> : (void*)dense2_chain_multi<true, 2>)
                   : (ns == 3 ? (void*)dense2_chain_multi<false, 3> : (void*)dense2_chain_multi<false, 2>);
    cudaLaunchCooperativeKernel(fn, dim3(q.grid), dim3(D2_THREADS), args, q.smemM, st);
}

// after dense_dual_setup (which decided DenseDev::nstrip and allocated the pre-formed generators): kernel attributes
// of the tiled multi-term chain; falls back to one term per barrier if the stages do not fit
inline int dense2_multi_setup(Dense2Plan& q, DensePlan& dp, std::string& err) {
    DenseDev& d = dp.d;
    if (!q.on || d.nstrip < 2) return 0;
    const size_t smem = sizeof(double) * D2_ST * (d.nstrip == 3 ? d2m_stage_doubles<3>() : d2m_stage_doubles<2>());
    cudaError_t e = cudaErrorInvalidValue;
    if (smem <= 227 * 1024) {
        if (d.nstrip == 3) {
            e = cudaFuncSetAttribute(dense2_chain_multi<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(dense2_chain_multi<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        } else {
            e = cudaFuncSetAttribute(dense2_chain_multi<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(dense2_chain_multi<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
    }
    if (e != cudaSuccess) { cudaGetLastError(); d.nstrip = 1; return 0; }
    q.smemM = smem;
    (void)err;
    return 0;
}
