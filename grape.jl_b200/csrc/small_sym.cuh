// small_sym.cuh -- real-symmetric generators on the time-segmented schedule of the N <= 3 path.
//
// Closed-system control problems in a rotating frame with real control fields have
// H0 and every control operator real and symmetric (README two-level system, the
// Lambda system of test/test_state_running_cost.jl:183-227).  Then A = +i dt H^dagger
// = i Hs with Hs = dt H REAL, and both steps of the schedule of small_seg.cuh are done
// in real matrix arithmetic on complex vectors:
//
//   A2' small_formseg_sym   U_n = exp(-i Hs) = cos(Hs) - i sin(Hs); cos and sin are
//        polynomials in the real matrix Q = Hs^2 (the even / odd halves of the same
//        degree-3/7/11/15 Taylor polynomial exp_plan() prescribes), a real N x N product
//        costs a quarter of a complex one; P_seg <- U_n P_seg is four real products.
//   C2' small_seggrad_sym   Krylov vectors of the END-of-step states (seg_step_krylov_h)
//        written with the un-normalised real-matrix powers  w_a = Hs^a Psi, x_b = Hs^b chi:
//          bt_a = i^a w_a / a!,   ch_b = i^b x_b / b!      (i^a = swap / negate, free)
//          Psi(t_{n-1}) = sum_a bt_a,  chi(t_{n-1}) = sum_b ch_b
//          M = sum_{a+b<=m-1} bt_a ch_b^dagger/(a+b+1) = sum_b e_b (i^b x_b)^dagger,
//          e_b = sum_a kappa(a,b) i^a w_a,  kappa(a,b) = 1/((a+b+1) a! b!)
//        and because E_l = i s_l mu_l^dagger has a REAL mu_l, the gradient element
//          Re tau_grad[k][n,l] = rho_k s_l sum_pq mu_l[p,q] Im M[p,q]     (optimize.jl:893-895, 574-584)
//        needs only Im M.  The Taylor order is the warp maximum, so the step is a
//        branch-free, fully unrolled template <M>.
//
// Same truncation as the m-term GradGenerator block recursion (a + b <= m - 1), hence the
// same numbers as small_seggrad<N, LC, true> to rounding (tests/test_gpu_parity_segmented.py).
// Steps that would need more than SEG_MMAX orders or sub-stepping are detected by A2' on the
// device (SegArgs::notfast); then C2' returns at once and the general Hermitian kernel, launched
// right behind it with run_if = notfast, does the work.
#pragma once
#include "small_seg.cuh"
#include "econ.cuh"
#include <algorithm>

// Series tables of the real-symmetric kernels, filled by sym_tables_upload() when a handle is set up:
//   * Taylor (GRAPE_B200_ECON=0): th[m] = the largest theta with theta^m / m! <= 2e-17 (vec_plan's criterion), g = 1;
//   * economised (default; econ.cuh): the Chebyshev-cut polynomial p_m of exp(-i x) on [-th[m], th[m]] in the monomial
//     basis -- one to two orders fewer for the same 1e-17 (C3: m = 6 instead of 7 in the gradient step, and a degree-6
//     propagator polynomial instead of degree 7 in the formation).  The generators of this file are real symmetric,
//     so the spectrum of Hs lies in [-||Hs||_1, ||Hs||_1].
// ginv[m][a] = g[m][a] / a!;  kap[m][a][b] = g[m][a+b+1] / ((a+b+1) a! b!): the derivative of p_m(X) = sum_j g_j X^j / j!
// is sum_j g_j / j! sum_{a+b=j-1} X^a dX X^b.
constexpr int SYM_NCLS = 6;   // propagator classes of the formation kernels: degree 3, 5, 6, 8, 12, 16
struct SymTab {
    double th[SEG_MMAX + 1];
    double ginv[SEG_MMAX + 1][SEG_MMAX + 1];
    double kap[SEG_MMAX + 1][SEG_MMAX + 1][SEG_MMAX];
    double fth[SYM_NCLS];        // radius of class c
    double cc[SYM_NCLS][9];      // cos(x)   ~ sum_i cc[i] x^(2i),  i <= degree / 2
    double sc[SYM_NCLS][9];      // sin(x)/x ~ sum_i sc[i] x^(2i),  i <= (degree - 1) / 2
};
__constant__ SymTab c_sym;
constexpr int SYM_CLS_DEG[SYM_NCLS] = {3, 5, 6, 8, 12, 16};

inline cudaError_t sym_tables_upload(bool econ) {
    // once per device and setting: a second handle with the same setting must not rewrite the table under the kernels
    // of the first one (constant memory is per device, shared by all handles of the process)
    static int uploaded[64];   // 0: nothing, 1: Taylor, 2: economised
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && uploaded[dev] == (econ ? 2 : 1)) return cudaSuccess;
    static SymTab tab;
    memset(&tab, 0, sizeof tab);
    const EconTab& et = econ_table();
    long double fact[ECON_MAXM + 2];
    fact[0] = 1.0L;
    for (int j = 1; j <= ECON_MAXM + 1; ++j) fact[j] = fact[j - 1] * j;
    for (int m = 2; m <= SEG_MMAX; ++m) {
        // Taylor radius: theta^m / m! <= 2e-17
        tab.th[m] = econ ? et.theta[m] : (double)powl(2e-17L * fact[m], 1.0L / m) * (1.0 - 1e-15);
        for (int a = 0; a <= m; ++a) tab.ginv[m][a] = (double)((econ ? (long double)et.g[m][a] : 1.0L) / fact[a]);
        for (int a = 0; a <= SEG_MMAX; ++a)
            for (int b = 0; b < SEG_MMAX; ++b) {
                const int j = a + b + 1;
                const long double g = (econ && j <= m) ? (long double)et.g[m][j] : 1.0L;
                tab.kap[m][a][b] = (double)(g / ((long double)j * fact[a] * fact[b]));
            }
    }
    for (int c = 0; c < SYM_NCLS; ++c) {
        const int m = SYM_CLS_DEG[c];
        // Taylor radius of degree m: first dropped term theta^(m+1) / (m+1)! <= 1e-17
        tab.fth[c] = econ ? et.theta[m] : std::min(1.0, (double)powl(1e-17L * fact[m + 1], 1.0L / (m + 1)) * (1.0 - 1e-15));
        for (int j = 0; j <= m; ++j) {
            const long double v = (econ ? (long double)et.g[m][j] : 1.0L) / fact[j] * (((j / 2) & 1) ? -1.0L : 1.0L);
            if (j & 1) tab.sc[c][j / 2] = (double)v; else tab.cc[c][j / 2] = (double)v;
        }
    }
    const cudaError_t e = cudaMemcpyToSymbol(c_sym, &tab, sizeof tab);
    if (e == cudaSuccess && dev >= 0 && dev < 64) uploaded[dev] = econ ? 2 : 1;
    return e;
}

// C = A*B, real N x N row-major in registers
template <int N>
GB_D void rm_mm(double (&C)[N * N], const double (&A)[N * N], const double (&B)[N * N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) acc = fma(A[i * N + k], B[k * N + j], acc);
            C[i * N + j] = acc;
        }
}
// y = A x, A real, x complex
template <int N>
GB_D void rm_mv(cplx (&y)[N], const double (&A)[N * N], const cplx (&x)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        cplx acc = mk(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < N; ++q) cfmar(acc, A[i * N + q], x[q]);
        y[i] = acc;
    }
}
// i^r w  (r = 0..3): swaps and sign flips only
GB_D cplx rot_i(cplx w, int r) {
    switch (r & 3) {
        case 0: return w;
        case 1: return mk(-w.y, w.x);
        case 2: return mk(-w.x, -w.y);
        default: return mk(w.y, -w.x);
    }
}

// Upper bound of the spectral radius of a real symmetric matrix: min(1-norm, Frobenius norm).  Neither dominates (a
// diagonal matrix: 1-norm exact, Frobenius up to sqrt(N) above; the Lambda system of C3: Frobenius 0.0065 where the
// 1-norm gives 0.0091, which decides between 5 and 6 orders per gradient step).
template <int N>
GB_D double sym_radius_bound(const double (&Hs)[N * N]) {
    double n1 = 0.0, f2 = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) { s += fabs(Hs[i * N + j]); f2 = fma(Hs[i * N + j], Hs[i * N + j], f2); }
        n1 = fmax(n1, s);
    }
    return fmin(n1, sqrt(f2) * (1.0 + 1e-14));
}

// Hs (unscaled: H0 + sum_l a_l Hc_l of generator g at step n) and the bound of its spectral radius
template <int N>
GB_D double sym_form_H(const DevP& p, const SegArgs& a, int g, int n, double (&Hs)[N * N]) {
    constexpr int NN = N * N;
    const int G = p.G, NT = p.NT;
#pragma unroll
    for (int c = 0; c < NN; ++c) Hs[c] = __ldg(&a.H0r[(size_t)c * G + g]);
    for (int l = 0; l < p.L; ++l) {
        double am = p.eps[l * NT + n];
        if (p.shape) am *= p.shape[l * NT + n];
#pragma unroll
        for (int c = 0; c < NN; ++c) Hs[c] = fma(am, __ldg(&a.Hcr[((size_t)l * NN + c) * G + g]), Hs[c]);
    }
    return sym_radius_bound<N>(Hs);
}

// ---------------------------------------------------------------------------
// A2': segment propagators, thread per (g, seg)
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) small_formseg_sym(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int G = p.G, NT = p.NT;
    if (idx >= (long long)G * a.NSEG) return;
    const int g = (int)(idx % G), seg = (int)(idx / G);
    const int n0 = seg * a.S, n1 = min(NT, n0 + a.S);
    double Pr[NN], Pi[NN];
    for (int n = n0; n < n1; ++n) {
        const double dt = p.tlist[n + 1] - p.tlist[n];
        double Hs[NN];
        const double theta = dt * sym_form_H<N>(p, a, g, n, Hs);
        if (theta > c_sym.th[SEG_MMAX]) *a.notfast = 1;   // the gradient kernel of this file cannot serve this step
        int degree, s;
        exp_plan(theta, degree, s);
        const double sc = s > 0 ? dt * ldexp(1.0, -s) : dt;
#pragma unroll
        for (int c = 0; c < NN; ++c) Hs[c] *= sc;
        double Q[NN];
        rm_mm<N>(Q, Hs, Hs);
        // cos(Hs) = sum_j (-1)^j Q^j/(2j)!,  sin(Hs) = Hs sum_j (-1)^j Q^j/(2j+1)!,  j <= d2 = (degree-1)/2
        const int d2 = (degree - 1) / 2;
        double Cm[NN], Sp[NN];
        {
            const double sg = (d2 & 1) ? -1.0 : 1.0;
            const double c1 = sg * c_invfact[2 * d2], c0 = -sg * c_invfact[2 * d2 - 2];
            const double s1 = sg * c_invfact[2 * d2 + 1], s0 = -sg * c_invfact[2 * d2 - 1];
#pragma unroll
            for (int c = 0; c < NN; ++c) { Cm[c] = c1 * Q[c]; Sp[c] = s1 * Q[c]; }
#pragma unroll
            for (int i = 0; i < N; ++i) { Cm[i * N + i] += c0; Sp[i * N + i] += s0; }
        }
        for (int j = d2 - 2; j >= 0; --j) {
            const double sg = (j & 1) ? -1.0 : 1.0;
            const double cj = sg * c_invfact[2 * j], sj = sg * c_invfact[2 * j + 1];
            double T1[NN], T2[NN];
            rm_mm<N>(T1, Q, Cm);
            rm_mm<N>(T2, Q, Sp);
#pragma unroll
            for (int c = 0; c < NN; ++c) { Cm[c] = T1[c]; Sp[c] = T2[c]; }
#pragma unroll
            for (int i = 0; i < N; ++i) { Cm[i * N + i] += cj; Sp[i * N + i] += sj; }
        }
        double Sm[NN];
        rm_mm<N>(Sm, Hs, Sp);
        for (int t = 0; t < s; ++t) {   // (C - iS)^2 = (C^2 - S^2) - i (2 S C)
            double T1[NN], T2[NN], T3[NN];
            rm_mm<N>(T1, Cm, Cm);
            rm_mm<N>(T2, Sm, Sm);
            rm_mm<N>(T3, Sm, Cm);
#pragma unroll
            for (int c = 0; c < NN; ++c) { Cm[c] = T1[c] - T2[c]; Sm[c] = 2.0 * T3[c]; }
        }
        if (n == n0) {
#pragma unroll
            for (int c = 0; c < NN; ++c) { Pr[c] = Cm[c]; Pi[c] = -Sm[c]; }
        } else {   // (C - iS)(Pr + i Pi)
            double Tr[NN], Ti[NN];
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double ar = 0.0, ai = 0.0;
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        ar = fma(Cm[i * N + k], Pr[k * N + j], ar);
                        ar = fma(Sm[i * N + k], Pi[k * N + j], ar);
                        ai = fma(Cm[i * N + k], Pi[k * N + j], ai);
                        ai = fma(-Sm[i * N + k], Pr[k * N + j], ai);
                    }
                    Tr[i * N + j] = ar;
                    Ti[i * N + j] = ai;
                }
#pragma unroll
            for (int c = 0; c < NN; ++c) { Pr[c] = Tr[c]; Pi[c] = Ti[c]; }
        }
    }
    cplx* o = a.Pseg + (size_t)seg * NN * G + g;
#pragma unroll
    for (int c = 0; c < NN; ++c) o[(size_t)c * G] = mk(Pr[c], Pi[c]);
}

// ---------------------------------------------------------------------------
// C2': one backward step with M Taylor orders. Hs = dt H (real). On exit psi, chi hold the
// start-of-step states and IM = Im sum_{a+b<=M-1} bt_a ch_b^dagger/(a+b+1).
// ---------------------------------------------------------------------------
template <int N, int M>
GB_D void sym_step(const double (&Hs)[N * N], cplx (&psi)[N], cplx (&chi)[N], double (&IM)[N * N]) {
    cplx e[M][N];
    cplx w[N], ap[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { w[i] = psi[i]; ap[i] = psi[i]; }
#pragma unroll
    for (int b = 0; b < M; ++b)
#pragma unroll
        for (int i = 0; i < N; ++i) e[b][i] = cscale(w[i], c_sym.kap[M][0][b]);
#pragma unroll
    for (int aa = 1; aa <= M; ++aa) {
        cplx nw[N];
        rm_mv<N>(nw, Hs, w);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            w[i] = nw[i];
            cfmar(ap[i], c_sym.ginv[M][aa], rot_i(nw[i], aa));
        }
#pragma unroll
        for (int b = 0; b < M - aa; ++b)
#pragma unroll
            for (int i = 0; i < N; ++i) cfmar(e[b][i], c_sym.kap[M][aa][b], rot_i(nw[i], aa));
    }
    cplx x[N], ac[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { x[i] = chi[i]; ac[i] = chi[i]; }
#pragma unroll
    for (int c = 0; c < N * N; ++c) IM[c] = 0.0;
#pragma unroll
    for (int b = 0; b < M; ++b) {
#pragma unroll
        for (int q = 0; q < N; ++q) {
            const cplx z = rot_i(x[q], b);   // Im(e conj(z)) = e.y z.x - e.x z.y
#pragma unroll
            for (int pp = 0; pp < N; ++pp) {
                double t = IM[pp * N + q];
                t = fma(e[b][pp].y, z.x, t);
                t = fma(-e[b][pp].x, z.y, t);
                IM[pp * N + q] = t;
            }
        }
        cplx nx[N];
        rm_mv<N>(nx, Hs, x);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            x[i] = nx[i];
            cfmar(ac[i], c_sym.ginv[M][b + 1], rot_i(nx[i], b + 1));
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) { psi[i] = ap[i]; chi[i] = ac[i]; }
}

// MINB = resident CTAs per SM the register allocation aims at (2: 255 registers, no spills; 3: 168 registers)
template <int N, int MINB>
__global__ void __launch_bounds__(128, MINB) small_seggrad_sym(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    if (*a.notfast) return;   // uniform over the grid: small_seggrad<N, LC, true> serves this call
    const int K = p.K, G = p.G, NT = p.NT;
    const int lane = threadIdx.x & 31;
    const int BKL = a.BKL, SPW = 32 / BKL;
    const int KGR = (K + BKL - 1) / BKL;
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int SGR = (a.NSEG + SPW - 1) / SPW;
    if (wid >= (long long)KGR * SGR) return;
    const int kg = (int)(wid % KGR), sg = (int)(wid / KGR);
    const int tk = lane % BKL, ts = lane / BKL;
    const int k = kg * BKL + tk, seg = sg * SPW + ts;
    const bool live = (k < K) && (seg < a.NSEG);
    const int kk = k < K ? k : K - 1;
    const int sseg = seg < a.NSEG ? seg : a.NSEG - 1;
    const int n0 = sseg * a.S, n1 = min(NT, n0 + a.S);
    const int g = p.gen[kk];
    const double rho = p.rho[kk];

    cplx chi[N], psi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        chi[i] = a.chiE[((size_t)sseg * N + i) * K + kk];
        psi[i] = ld_cs(&p.psi[((size_t)n1 * N + i) * K + kk]);   // segment boundary written by the forward chain
    }
    for (int st = 0; st < a.S; ++st) {   // uniform trip count across the warp
        const int n = n1 - 1 - st;
        const bool act = live && n >= n0;
        const int nn = n >= n0 ? n : n0;
        const double dt = p.tlist[nn + 1] - p.tlist[nn];
        double Hs[NN];
        const double theta = dt * sym_form_H<N>(p, a, g, nn, Hs);
        int m = 2;
#pragma unroll
        for (int j = 2; j < SEG_MMAX; ++j) m = theta > c_sym.th[j] ? j + 1 : m;
        m = __reduce_max_sync(0xffffffffu, m);   // more orders never hurt: one uniform branch per warp
#pragma unroll
        for (int c = 0; c < NN; ++c) Hs[c] *= dt;
        double IM[NN];
        switch (m) {
            case 2: sym_step<N, 2>(Hs, psi, chi, IM); break;
            case 3: sym_step<N, 3>(Hs, psi, chi, IM); break;
            case 4: sym_step<N, 4>(Hs, psi, chi, IM); break;
            case 5: sym_step<N, 5>(Hs, psi, chi, IM); break;
            case 6: sym_step<N, 6>(Hs, psi, chi, IM); break;
            case 7: sym_step<N, 7>(Hs, psi, chi, IM); break;
            default: sym_step<N, 8>(Hs, psi, chi, IM); break;
        }
        for (int l = 0; l < p.L; ++l) {
            double sl = dt * rho;
            if (p.dshape) sl *= p.dshape[l * NT + nn];
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < NN; ++c) acc = fma(__ldg(&a.Hcr[((size_t)l * NN + c) * G + g]), IM[c], acc);
            double red = act ? sl * acc : 0.0;
            // fixed-order sum over the BKL trajectories of this lane group
            for (int off = BKL >> 1; off > 0; off >>= 1) red += __shfl_down_sync(0xffffffffu, red, off, BKL);
            if (tk == 0 && seg < a.NSEG && n >= n0) p.partial[(size_t)kg * p.L * NT + (size_t)l * NT + n] = red;
        }
    }
}

// ===========================================================================================
// Round 2: the same two kernels with the thread's generator staged ONCE in shared memory.
//
// ncu / SASS of the kernels above (profiles/r1_s11_ncu_full_c3_sym_raw.csv): FP64 pipe 54-57 % active, but
// only ~60 % of the issued instructions are DFMA/DMUL -- every step re-loads the 9 (1 + L) real operator
// elements of the thread's generator through 64-bit address arithmetic (IMAD/LEA/IADD3 + LDG: ~170
// instructions per control and step, twice in the gradient kernel: Hs formation and the traces).  The
// operators do not depend on the step: each thread copies them once into its own column of a shared-memory
// tile sH[c][tid] (conflict-free, addressed by immediates), and the number of controls is a template
// parameter (LT = 1, 2; 0 = run-time L) so the per-control loops unroll.  Same arithmetic, same order of
// operations per element, hence bit-identical results to the kernels above.
// ===========================================================================================
constexpr int SYM_BD = 128;   // threads per block of the staged kernels

// dynamic shared memory of the staged kernels, bytes
inline size_t sym_stage_bytes(int N, int L) { return (size_t)(1 + L) * N * N * SYM_BD * sizeof(double); }

template <int N, int BD = SYM_BD>
GB_D void sym_stage(const DevP& p, const SegArgs& a, int g, int L, double* sH) {
    constexpr int NN = N * N;
    const int G = p.G, t = threadIdx.x;
#pragma unroll
    for (int c = 0; c < NN; ++c) sH[c * BD + t] = __ldg(&a.H0r[(size_t)c * G + g]);
    for (int l = 0; l < L; ++l)
#pragma unroll
        for (int c = 0; c < NN; ++c) sH[(NN + l * NN + c) * BD + t] = __ldg(&a.Hcr[((size_t)l * NN + c) * G + g]);
}

// Hs = H0 + sum_l a_l Hc_l from the staged tile, and its 1-norm
// amplitude of control l at step n (eps * shape; in amplitude mode eps holds the amplitude itself and shape is off)
GB_D double sym_amp(const DevP& p, int l, int n) {
    double am = p.eps[l * p.NT + n];
    if (p.shape) am *= p.shape[l * p.NT + n];
    return am;
}
// Hs = H0 + sum_l am[l] Hc_l from the staged tile with the amplitudes already in registers, and its 1-norm
template <int N, int LT, int BD = SYM_BD>
GB_D double sym_form_H_amps(const double* sH, const double (&am)[LT > 0 ? LT : 1], double (&Hs)[N * N]) {
    constexpr int NN = N * N;
    const int t = threadIdx.x;
#pragma unroll
    for (int c = 0; c < NN; ++c) Hs[c] = sH[c * BD + t];
#pragma unroll
    for (int l = 0; l < LT; ++l)
#pragma unroll
        for (int c = 0; c < NN; ++c) Hs[c] = fma(am[l], sH[(NN + l * NN + c) * BD + t], Hs[c]);
    return sym_radius_bound<N>(Hs);
}
template <int N, int LT, int BD = SYM_BD>
GB_D double sym_form_H_staged(const DevP& p, const double* sH, int L, int n, double (&Hs)[N * N]) {
    constexpr int NN = N * N;
    const int NT = p.NT, t = threadIdx.x;
#pragma unroll
    for (int c = 0; c < NN; ++c) Hs[c] = sH[c * BD + t];
    if (LT > 0) {
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            double am = p.eps[l * NT + n];
            if (p.shape) am *= p.shape[l * NT + n];
#pragma unroll
            for (int c = 0; c < NN; ++c) Hs[c] = fma(am, sH[(NN + l * NN + c) * BD + t], Hs[c]);
        }
    } else {
        for (int l = 0; l < L; ++l) {
            double am = p.eps[l * NT + n];
            if (p.shape) am *= p.shape[l * NT + n];
#pragma unroll
            for (int c = 0; c < NN; ++c) Hs[c] = fma(am, sH[(NN + l * NN + c) * BD + t], Hs[c]);
        }
    }
    return sym_radius_bound<N>(Hs);
}

// C = A B for two COMMUTING symmetric matrices (polynomials in the same Hs): the product is symmetric, only the upper
// triangle is computed (N = 3: 6 dot products instead of 9) and both halves hold the same rounded value
template <int N>
GB_D void rm_mm_sym(double (&C)[N * N], const double (&A)[N * N], const double (&B)[N * N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) acc = fma(A[i * N + k], B[k * N + j], acc);
            C[i * N + j] = acc;
            C[j * N + i] = acc;
        }
}

// Cm = cos(Hs) and Sp = sin(Hs)/Hs as Horner polynomials in Q = Hs^2: the even / odd halves of the degree-DEG
// polynomial of class CLS (c_sym: Chebyshev-cut or Taylor).  CLS is a template parameter, so the coefficients are
// constant-bank operands and the loops unroll (a run-time loop cost 15 % of the formation kernel's instructions).
template <int N, int CLS>
GB_D void sym_cos_sinc_poly(const double (&Q)[N * N], double (&Cm)[N * N], double (&Sp)[N * N]) {
    constexpr int NN = N * N;
    constexpr int DEG = SYM_CLS_DEG[CLS], DC = DEG / 2, DS = (DEG - 1) / 2;   // DC = DS (odd degree) or DS + 1 (even)
    static_assert(DS >= 1 && (DC == DS || DC == DS + 1), "degree >= 3");
#pragma unroll
    for (int c = 0; c < NN; ++c) { Cm[c] = c_sym.cc[CLS][DC] * Q[c]; Sp[c] = c_sym.sc[CLS][DS] * Q[c]; }
#pragma unroll
    for (int i = 0; i < N; ++i) { Cm[i * N + i] += c_sym.cc[CLS][DC - 1]; Sp[i * N + i] += c_sym.sc[CLS][DS - 1]; }
    if (DC > DS) {   // even degree: the cosine has one term more -- one product on its own, then both series in step
        double T1[NN];
        rm_mm_sym<N>(T1, Q, Cm);
#pragma unroll
        for (int c = 0; c < NN; ++c) Cm[c] = T1[c];
#pragma unroll
        for (int i = 0; i < N; ++i) Cm[i * N + i] += c_sym.cc[CLS][DC - 2];
    }
#pragma unroll
    for (int j = DS - 2; j >= 0; --j) {
        double T1[NN], T2[NN];
        rm_mm_sym<N>(T1, Q, Cm);
        rm_mm_sym<N>(T2, Q, Sp);
#pragma unroll
        for (int c = 0; c < NN; ++c) { Cm[c] = T1[c]; Sp[c] = T2[c]; }
#pragma unroll
        for (int i = 0; i < N; ++i) { Cm[i * N + i] += c_sym.cc[CLS][j]; Sp[i * N + i] += c_sym.sc[CLS][j]; }
    }
}

// U_n = cos(Hs) - i sin(Hs) of one step (Hs unscaled generator, theta >= dt * its spectral radius: sym_radius_bound)
template <int N>
GB_D void sym_cos_sin(double (&Hs)[N * N], double dt, double theta, double (&Cm)[N * N], double (&Sm)[N * N]) {
    constexpr int NN = N * N;
    int cls = 0, s = 0;
#pragma unroll
    for (int c = 0; c < SYM_NCLS - 1; ++c) cls = theta > c_sym.fth[c] ? c + 1 : cls;
    {
        double t = theta;
        while (t > c_sym.fth[SYM_NCLS - 1] && s < 60) { t *= 0.5; ++s; }   // scaling and squaring beyond the top class
    }
    const double sc = s > 0 ? dt * ldexp(1.0, -s) : dt;
#pragma unroll
    for (int c = 0; c < NN; ++c) Hs[c] *= sc;
    double Q[NN];
    rm_mm_sym<N>(Q, Hs, Hs);
    double Sp[NN];
    switch (cls) {
        case 0: sym_cos_sinc_poly<N, 0>(Q, Cm, Sp); break;
        case 1: sym_cos_sinc_poly<N, 1>(Q, Cm, Sp); break;
        case 2: sym_cos_sinc_poly<N, 2>(Q, Cm, Sp); break;
        case 3: sym_cos_sinc_poly<N, 3>(Q, Cm, Sp); break;
        case 4: sym_cos_sinc_poly<N, 4>(Q, Cm, Sp); break;
        default: sym_cos_sinc_poly<N, 5>(Q, Cm, Sp); break;
    }
    rm_mm_sym<N>(Sm, Hs, Sp);
    for (int t = 0; t < s; ++t) {   // (C - iS)^2 = (C^2 - S^2) - i (2 S C)
        double T1[NN], T2[NN], T3[NN];
        rm_mm_sym<N>(T1, Cm, Cm);
        rm_mm_sym<N>(T2, Sm, Sm);
        rm_mm_sym<N>(T3, Sm, Cm);
#pragma unroll
        for (int c = 0; c < NN; ++c) { Cm[c] = T1[c] - T2[c]; Sm[c] = 2.0 * T3[c]; }
    }
}

// B <- (C - iS) A on split real / imaginary arrays (four real products)
template <int N>
GB_D void sym_apply_U(const double (&Cm)[N * N], const double (&Sm)[N * N], const double (&Ar)[N * N], const double (&Ai)[N * N],
                      double (&Br)[N * N], double (&Bi)[N * N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double ar = 0.0, ai = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                ar = fma(Cm[i * N + k], Ar[k * N + j], ar);
                ar = fma(Sm[i * N + k], Ai[k * N + j], ar);
                ai = fma(Cm[i * N + k], Ai[k * N + j], ai);
                ai = fma(-Sm[i * N + k], Ar[k * N + j], ai);
            }
            Br[i * N + j] = ar;
            Bi[i * N + j] = ai;
        }
}

template <int N, int LT, int MINB = 3>
__global__ void __launch_bounds__(SYM_BD, MINB) small_formseg_sym2(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    extern __shared__ double sH[];
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int G = p.G, NT = p.NT, L = LT > 0 ? LT : p.L;
    if (idx >= (long long)G * a.NSEG) return;
    const int g = (int)(idx % G), seg = (int)(idx / G);
    const int n0 = seg * a.S, n1 = min(NT, n0 + a.S);
    sym_stage<N>(p, a, g, L, sH);
    double Pr[NN], Pi[NN];
    auto step_U = [&](int n, double (&Cm)[NN], double (&Sm)[NN]) {
        const double dt = p.tlist[n + 1] - p.tlist[n];
        double Hs[NN];
        const double theta = dt * sym_form_H_staged<N, LT>(p, sH, L, n, Hs);
        sym_cos_sin<N>(Hs, dt, theta, Cm, Sm);
    };
    {
        double Cm[NN], Sm[NN];
        step_U(n0, Cm, Sm);
#pragma unroll
        for (int c = 0; c < NN; ++c) { Pr[c] = Cm[c]; Pi[c] = -Sm[c]; }
    }
    // two steps per trip, P -> T -> P: the product lands where the next step reads it (a one-step loop copied the 2 N^2
    // doubles back every step: 36 of the 430 instructions per step were register moves, ncu source page r2_s19)
    for (int n = n0 + 1; n < n1; n += 2) {
        double Tr[NN], Ti[NN];
        {
            double Cm[NN], Sm[NN];
            step_U(n, Cm, Sm);
            sym_apply_U<N>(Cm, Sm, Pr, Pi, Tr, Ti);
        }
        if (n + 1 < n1) {
            double Cm[NN], Sm[NN];
            step_U(n + 1, Cm, Sm);
            sym_apply_U<N>(Cm, Sm, Tr, Ti, Pr, Pi);
        } else {
#pragma unroll
            for (int c = 0; c < NN; ++c) { Pr[c] = Tr[c]; Pi[c] = Ti[c]; }
        }
    }
    cplx* o = a.Pseg + (size_t)seg * NN * G + g;
#pragma unroll
    for (int c = 0; c < NN; ++c) o[(size_t)c * G] = mk(Pr[c], Pi[c]);
}

// rare steps beyond 8 orders: nsub equal sub-steps (kept out of line so that the hot path's registers are not affected)
template <int N>
__device__ __noinline__ void sym_step_sub(double (&Hs)[N * N], cplx (&psi)[N], cplx (&chi)[N], double (&IM)[N * N], int nsub) {
    constexpr int NN = N * N;
    const double f = 1.0 / (double)nsub;
#pragma unroll
    for (int c = 0; c < NN; ++c) { Hs[c] *= f; IM[c] = 0.0; }
    for (int sub = 0; sub < nsub; ++sub) {
        double IS[NN];
        sym_step<N, 8>(Hs, psi, chi, IS);
#pragma unroll
        for (int c = 0; c < NN; ++c) IM[c] = fma(f, IS[c], IM[c]);
    }
}

// orders above the inline limit MI of the gradient kernel (out of line like the sub-stepped steps: the register budget
// of the kernel is then set by sym_step<N, MI>, not by the rarely needed sym_step<N, 8>)
template <int N>
__device__ __noinline__ void sym_step_hi(const double (&Hs)[N * N], cplx (&psi)[N], cplx (&chi)[N], double (&IM)[N * N], int m) {
    if (m <= 7) sym_step<N, 7>(Hs, psi, chi, IM);
    else sym_step<N, 8>(Hs, psi, chi, IM);
}

// BD / MINB / MI: 128 threads; launched with MINB = 4 (128 registers, 16 warps / SM) and MI = 6 (orders 7, 8 out of
// line) for one or two controls, MINB = 3 (168 registers) and all orders inline for a run-time number of controls.
// Measured on C3 (profiles/r2_s18_c3_variants.txt, r2_s19_c3_variants.txt): 128 registers with all orders inline is 8 %
// slower than 168 (spills), with MI = 6 it is 3 % faster; two-warp blocks (BD = 64) and the pulse-value prefetch (PF)
// gave nothing; a 128-register build of the formation kernel is 3 % slower.
template <int N, int LT, int MINB = 3, int BD = SYM_BD, bool PF = false, int MI = 8>
__global__ void __launch_bounds__(BD, MINB) small_seggrad_sym2(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    extern __shared__ double sH[];
    const int K = p.K, NT = p.NT, L = LT > 0 ? LT : p.L;
    const int lane = threadIdx.x & 31;
    const int BKL = a.BKL, SPW = 32 / BKL;
    const int KGR = (K + BKL - 1) / BKL;
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int SGR = (a.NSEG + SPW - 1) / SPW;
    if (wid >= (long long)KGR * SGR) return;
    const int kg = (int)(wid % KGR), sg = (int)(wid / KGR);
    const int tk = lane % BKL, ts = lane / BKL;
    const int k = kg * BKL + tk, seg = sg * SPW + ts;
    const bool live = (k < K) && (seg < a.NSEG);
    const int kk = k < K ? k : K - 1;
    const int sseg = seg < a.NSEG ? seg : a.NSEG - 1;
    const int n0 = sseg * a.S, n1 = min(NT, n0 + a.S);
    double rho = a.scan ? 1.0 : p.rho[kk];   // scan schedules: computed in the prologue below
    sym_stage<N, BD>(p, a, p.gen[kk], L, sH);

    cplx chi[N], psi[N];
    if (a.scan == 2) {
        // one-pass boundary chains (small_segchain_dual): chiE holds the propagated raw target, chi = (c_k / rho_k) x it
        const double w = p.w ? p.w[kk] : 1.0;
        const double Kg = (double)p.Kglobal;
        cplx c;
        if (p.functional == 0) c = mk(w * p.sums[0] / (Kg * Kg), w * p.sums[1] / (Kg * Kg));
        else if (p.functional == 1) c = mk(w / (2.0 * Kg), 0.0);
        else { cplx t = p.tau[kk]; c = mk(w * t.x / Kg, w * t.y / Kg); }
        cplx x[N];
        double rn = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) { x[i] = cmul(c, p.tgt[(size_t)i * K + kk]); rn += cnorm2(x[i]); }
        rn = sqrt(rn);
        if (!(rn >= p.chi_min_norm)) {   // optimize.jl:1021-1025
            if (live && seg == 0 && atomicCAS(&p.flags->chi_bad_k, 0, k + 1) == 0) p.flags->chi_bad_rho = rn;
            rn = 1.0;
        }
        rho = rn;
        const double ir = 1.0 / rn;
        if (live && seg == 0) {
            p.rho[k] = rn;
#pragma unroll
            for (int i = 0; i < N; ++i) p.chiT[(size_t)k * N + i] = cscale(x[i], ir);
        }
        const cplx f = cscale(c, ir);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            chi[i] = cmul(f, a.chiE[((size_t)sseg * N + i) * K + kk]);
            psi[i] = ld_cs(&p.psi[((size_t)n1 * N + i) * K + kk]);
        }
    } else if (a.scan) {
        // scan schedule: both boundary states of this (k, seg) straight from the prefix products Q (4 mat-vecs):
        //   Psi(end of seg) = Q_seg Psi(0),  chi(end of seg) = Q_seg Q_last^dagger chi(T)   (every P is unitary)
        const int G = p.G, g = p.gen[kk];
        cplx Q[NN], x[N];
#pragma unroll
        for (int c = 0; c < NN; ++c) Q[c] = __ldg(&a.Pseg[((size_t)sseg * NN + c) * G + g]);
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = p.psi0[(size_t)i * K + kk];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j) cfma(acc, Q[i * N + j], x[j]);
            psi[i] = acc;
        }
        if (a.chi_host) {
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = a.chi_host[(size_t)kk * N + i];
        } else {   // optimize.jl:845-855 with the analytic chi of J_T_sm / J_T_re / J_T_ss
            const double w = p.w ? p.w[kk] : 1.0;
            const double Kg = (double)p.Kglobal;
            cplx c;
            if (p.functional == 0) c = mk(w * p.sums[0] / (Kg * Kg), w * p.sums[1] / (Kg * Kg));
            else if (p.functional == 1) c = mk(w / (2.0 * Kg), 0.0);
            else { cplx t = p.tau[kk]; c = mk(w * t.x / Kg, w * t.y / Kg); }
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = cmul(c, p.tgt[(size_t)i * K + kk]);
        }
        double rn = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) rn += cnorm2(x[i]);
        rn = sqrt(rn);
        if (!(rn >= p.chi_min_norm)) {   // optimize.jl:1021-1025
            if (live && seg == 0 && atomicCAS(&p.flags->chi_bad_k, 0, k + 1) == 0) p.flags->chi_bad_rho = rn;
            rn = 1.0;
        }
        rho = rn;
        const double ir = 1.0 / rn;
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = cscale(x[i], ir);
        if (live && seg == 0) {
            p.rho[k] = rn;
#pragma unroll
            for (int i = 0; i < N; ++i) p.chiT[(size_t)k * N + i] = x[i];
        }
        if (sseg < a.NSEG - 1) {
            cplx y[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < N; ++j)
                    cfmac(acc, __ldg(&a.Pseg[((size_t)(a.NSEG - 1) * NN + j * N + i) * G + g]), x[j]);   // (Q_last^dagger x)_i
                y[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < N; ++j) cfma(acc, Q[i * N + j], y[j]);
                chi[i] = acc;
            }
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) chi[i] = x[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            chi[i] = a.chiE[((size_t)sseg * N + i) * K + kk];
            psi[i] = ld_cs(&p.psi[((size_t)n1 * N + i) * K + kk]);   // segment boundary written by the forward chain
        }
    }
    double* const part = p.partial + (size_t)kg * L * NT;
    const bool writer = tk == 0 && seg < a.NSEG;
    // PF: the pulse values and the step width of the NEXT step are fetched while the current step computes.  Measured:
    // no gain (profiles/r2_s14_c3_ab.txt: 0.3789 ms with and without), so the default is the plain form.
    double am_next[LT > 0 ? LT : 1], dt_next;
    {
        const int nf = n1 - 1 >= n0 ? n1 - 1 : n0;
        dt_next = p.tlist[nf + 1] - p.tlist[nf];
#pragma unroll
        for (int l = 0; l < LT; ++l) am_next[l] = sym_amp(p, l, nf);
    }
    for (int st = 0; st < a.S; ++st) {   // uniform trip count across the warp
        const int n = n1 - 1 - st;
        const bool act = live && n >= n0;
        const int nn = n >= n0 ? n : n0;
        const double dt = dt_next;
        double Hs[NN];
        double theta;
        if (LT > 0 && PF) {
            double am[LT > 0 ? LT : 1];
#pragma unroll
            for (int l = 0; l < LT; ++l) am[l] = am_next[l];
            const int nx = n - 1 >= n0 ? n - 1 : n0;
            dt_next = p.tlist[nx + 1] - p.tlist[nx];
#pragma unroll
            for (int l = 0; l < LT; ++l) am_next[l] = sym_amp(p, l, nx);
            theta = dt * sym_form_H_amps<N, LT, BD>(sH, am, Hs);
        } else {
            const int nx = n - 1 >= n0 ? n - 1 : n0;
            dt_next = p.tlist[nx + 1] - p.tlist[nx];
            theta = dt * sym_form_H_staged<N, LT, BD>(p, sH, L, nn, Hs);
        }
        int m = 2;
#pragma unroll
        for (int j = 2; j < SEG_MMAX; ++j) m = theta > c_sym.th[j] ? j + 1 : m;
        // steps beyond 8 orders (||H dt||_1 > 0.0308): nsub equal sub-steps of 8 orders each -- the integral
        // M = int_0^1 Psi(s) chi(s)^dagger ds of the step is the mean of the sub-steps' integrals.  Decided per warp and
        // per step (round 1 sent the WHOLE call to the 1.8x slower complex kernel if one step anywhere was ineligible).
        int nsub = theta > c_sym.th[SEG_MMAX] ? (int)ceil(theta / c_sym.th[SEG_MMAX]) : 1;
        m = __reduce_max_sync(0xffffffffu, m);   // more orders never hurt: one uniform branch per warp
        nsub = __reduce_max_sync(0xffffffffu, nsub);
#pragma unroll
        for (int c = 0; c < NN; ++c) Hs[c] *= dt;
        double IM[NN];
        if (nsub == 1 && m <= MI) {
            switch (m) {
                case 2: sym_step<N, 2>(Hs, psi, chi, IM); break;
                case 3: sym_step<N, 3>(Hs, psi, chi, IM); break;
                case 4: sym_step<N, 4>(Hs, psi, chi, IM); break;
                case 5: sym_step<N, 5>(Hs, psi, chi, IM); break;
                case 6: sym_step<N, MI >= 6 ? 6 : 2>(Hs, psi, chi, IM); break;
                case 7: sym_step<N, MI >= 7 ? 7 : 2>(Hs, psi, chi, IM); break;
                default: sym_step<N, MI >= 8 ? 8 : 2>(Hs, psi, chi, IM); break;
            }
        } else if (nsub == 1) {   // m > MI
            double Hs2[NN], IM2[NN];
            cplx psi2[N], chi2[N];
#pragma unroll
            for (int c = 0; c < NN; ++c) Hs2[c] = Hs[c];
#pragma unroll
            for (int i = 0; i < N; ++i) { psi2[i] = psi[i]; chi2[i] = chi[i]; }
            sym_step_hi<N>(Hs2, psi2, chi2, IM2, m);
#pragma unroll
            for (int c = 0; c < NN; ++c) IM[c] = IM2[c];
#pragma unroll
            for (int i = 0; i < N; ++i) { psi[i] = psi2[i]; chi[i] = chi2[i]; }
        } else {   // copies: only they live in local memory for the out-of-line call
            double Hs2[NN], IM2[NN];
            cplx psi2[N], chi2[N];
#pragma unroll
            for (int c = 0; c < NN; ++c) Hs2[c] = Hs[c];
#pragma unroll
            for (int i = 0; i < N; ++i) { psi2[i] = psi[i]; chi2[i] = chi[i]; }
            sym_step_sub<N>(Hs2, psi2, chi2, IM2, nsub);
#pragma unroll
            for (int c = 0; c < NN; ++c) IM[c] = IM2[c];
#pragma unroll
            for (int i = 0; i < N; ++i) { psi[i] = psi2[i]; chi[i] = chi2[i]; }
        }
        // The control operators are read from shared memory twice per step (H formation above, the traces below).  Without
        // the volatile read the compiler keeps the 9 L values of the first read alive across sym_step -- in local memory:
        // 20 STL.64 + 20 LDL.64 per step (ncu source page, profiles/r2_s18_ncu_c3_source.csv) -- instead of re-reading.
        const volatile double* const sHv = sH;
        auto trace = [&](int l) {
            double sl = dt * rho;
            if (p.dshape) sl *= p.dshape[l * NT + nn];
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < NN; ++c) acc = fma(sHv[(NN + l * NN + c) * BD + threadIdx.x], IM[c], acc);
            return act ? sl * acc : 0.0;
        };
        if (LT > 0) {
            // the LT fixed-order sums over the BKL trajectories of this lane group run interleaved: one chain of
            // dependent SHFL + DADD per control was the kernel's largest scoreboard stall
            double red[LT > 0 ? LT : 1];
#pragma unroll
            for (int l = 0; l < LT; ++l) red[l] = trace(l);
            for (int off = BKL >> 1; off > 0; off >>= 1) {
#pragma unroll
                for (int l = 0; l < LT; ++l) red[l] += __shfl_down_sync(0xffffffffu, red[l], off, BKL);
            }
            if (writer && n >= n0) {
#pragma unroll
                for (int l = 0; l < LT; ++l) part[(size_t)l * NT + n] = red[l];
            }
        } else {
            for (int l = 0; l < L; ++l) {
                double red = trace(l);
                for (int off = BKL >> 1; off > 0; off >>= 1) red += __shfl_down_sync(0xffffffffu, red, off, BKL);
                if (writer && n >= n0) part[(size_t)l * NT + n] = red;
            }
        }
    }
}

// ===========================================================================================
// Round 2: no sequential chains at all -- the segment propagators are combined by a PARALLEL SCAN.
//
// B1 / B2 (small_segchain_fwd / _bwd) walk the NSEG segment boundaries one after the other: 18 + 13 us of pure
// dependency latency per gradient at K = 4096 and 23 + 21 us of a 130 us step at K = 512 (a shard of the ensemble
// on 8 GPUs; profiles/r2_s1_c3_sweep.txt).  Matrix products are associative, so the prefix products
//     Q_seg = P_seg P_{seg-1} ... P_0
// come from a Kogge-Stone scan across the lanes of the warp that formed the P_seg of one generator (5 levels of
// 3 x 3 complex products exchanged by shuffles, +4 % work in the formation kernel), and with them every boundary
// state is ONE mat-vec, independent of all the others:
//     Psi_k(end of seg) = Q_seg Psi_k(0),      tau_k = <tgt_k | Q_last Psi_k(0)>,
//     chi_k(end of seg) = (P_last .. P_{seg+1})^dagger chi_k(T) = Q_seg Q_last^dagger chi_k(T)   (Hermitian generators:
//     every P is unitary).
//   A2'' small_formscan_sym      one block per generator, thread = segment: P_seg as in A2', then the scan
//   B'   small_scan_tau_reduce   tau_k, final states and the sums of reduce_tau (single block, fixed order)
//   B''  small_scan_bounds       thread per (k, seg): chi_k(T) (optimize.jl:845-869), Psi / chi at the segment ends
// The gradient kernel C2' is unchanged (it reads the same boundary arrays the chains used to fill).
// ===========================================================================================

// C = A B, complex N x N on split real / imaginary arrays
template <int N>
GB_D void cm_mm(double (&Cr)[N * N], double (&Ci)[N * N], const double (&Ar)[N * N], const double (&Ai)[N * N],
                const double (&Br)[N * N], const double (&Bi)[N * N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double ar = 0.0, ai = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                ar = fma(Ar[i * N + k], Br[k * N + j], ar);
                ar = fma(-Ai[i * N + k], Bi[k * N + j], ar);
                ai = fma(Ar[i * N + k], Bi[k * N + j], ai);
                ai = fma(Ai[i * N + k], Br[k * N + j], ai);
            }
            Cr[i * N + j] = ar;
            Ci[i * N + j] = ai;
        }
}

constexpr int SCAN_MAXW = 4;   // warps per generator block: NSEG <= 128

template <int N, int LT>
__global__ void __launch_bounds__(32 * SCAN_MAXW, 3) small_formscan_sym(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    __shared__ double sG[(1 + (LT > 0 ? LT : 16)) * NN];          // the block's generator (run-time L <= 16)
    __shared__ double sT[SCAN_MAXW][2 * NN];                       // warp totals
    const int g = blockIdx.x, seg = threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int G = p.G, NT = p.NT, L = LT > 0 ? LT : p.L;
    for (int e = threadIdx.x; e < (1 + L) * NN; e += blockDim.x)
        sG[e] = e < NN ? __ldg(&a.H0r[(size_t)e * G + g]) : __ldg(&a.Hcr[(size_t)(e - NN) * G + g]);
    __syncthreads();
    const bool valid = seg < a.NSEG;
    const int n0 = seg * a.S, n1 = valid ? min(NT, n0 + a.S) : n0;
    double Pr[NN], Pi[NN];
#pragma unroll
    for (int c = 0; c < NN; ++c) { Pr[c] = 0.0; Pi[c] = 0.0; }
#pragma unroll
    for (int i = 0; i < N; ++i) Pr[i * N + i] = 1.0;               // lanes past the last segment carry the identity
    auto step_U = [&](int n, double (&Cm)[NN], double (&Sm)[NN]) {
        const double dt = p.tlist[n + 1] - p.tlist[n];
        double Hs[NN];
#pragma unroll
        for (int c = 0; c < NN; ++c) Hs[c] = sG[c];
        auto add_control = [&](int l) {
            double am = p.eps[l * NT + n];
            if (p.shape) am *= p.shape[l * NT + n];
#pragma unroll
            for (int c = 0; c < NN; ++c) Hs[c] = fma(am, sG[NN + l * NN + c], Hs[c]);
        };
        if (LT > 0) {
#pragma unroll
            for (int l = 0; l < LT; ++l) add_control(l);
        } else {
            for (int l = 0; l < L; ++l) add_control(l);
        }
        const double nrm = sym_radius_bound<N>(Hs);
        sym_cos_sin<N>(Hs, dt, dt * nrm, Cm, Sm);
    };
    if (n0 < n1) {
        double Cm[NN], Sm[NN];
        step_U(n0, Cm, Sm);
#pragma unroll
        for (int c = 0; c < NN; ++c) { Pr[c] = Cm[c]; Pi[c] = -Sm[c]; }
    }
    for (int n = n0 + 1; n < n1; n += 2) {   // two steps per trip, P -> T -> P (small_formseg_sym2)
        double Tr[NN], Ti[NN];
        {
            double Cm[NN], Sm[NN];
            step_U(n, Cm, Sm);
            sym_apply_U<N>(Cm, Sm, Pr, Pi, Tr, Ti);
        }
        if (n + 1 < n1) {
            double Cm[NN], Sm[NN];
            step_U(n + 1, Cm, Sm);
            sym_apply_U<N>(Cm, Sm, Tr, Ti, Pr, Pi);
        } else {
#pragma unroll
            for (int c = 0; c < NN; ++c) { Pr[c] = Tr[c]; Pi[c] = Ti[c]; }
        }
    }
    // inclusive scan over the lanes: lane i ends with P_i P_{i-1} .. P_{first lane of the warp}
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        double Tr[NN], Ti[NN];
#pragma unroll
        for (int c = 0; c < NN; ++c) {
            Tr[c] = __shfl_up_sync(0xffffffffu, Pr[c], off);
            Ti[c] = __shfl_up_sync(0xffffffffu, Pi[c], off);
        }
        if (lane >= off) {
            double Nr[NN], Ni[NN];
            cm_mm<N>(Nr, Ni, Pr, Pi, Tr, Ti);
#pragma unroll
            for (int c = 0; c < NN; ++c) { Pr[c] = Nr[c]; Pi[c] = Ni[c]; }
        }
    }
    if (blockDim.x > 32) {   // earlier warps of the same generator
        if (lane == 31) {
#pragma unroll
            for (int c = 0; c < NN; ++c) { sT[w][c] = Pr[c]; sT[w][NN + c] = Pi[c]; }
        }
        __syncthreads();
        for (int ww = w - 1; ww >= 0; --ww) {   // Q <- Q T_{w-1} T_{w-2} .. T_0
            double Tr[NN], Ti[NN], Nr[NN], Ni[NN];
#pragma unroll
            for (int c = 0; c < NN; ++c) { Tr[c] = sT[ww][c]; Ti[c] = sT[ww][NN + c]; }
            cm_mm<N>(Nr, Ni, Pr, Pi, Tr, Ti);
#pragma unroll
            for (int c = 0; c < NN; ++c) { Pr[c] = Nr[c]; Pi[c] = Ni[c]; }
        }
    }
    if (valid) {
        cplx* o = a.Pseg + (size_t)seg * NN * G + g;
#pragma unroll
        for (int c = 0; c < NN; ++c) o[(size_t)c * G] = mk(Pr[c], Pi[c]);
    }
}

// tau_k = <tgt_k | Q_last Psi_k(0)>  (optimize.jl:752-753), final states, and the sums of reduce_tau: thread per
// trajectory, one partial sum per block, the last block to arrive adds the partials in block order (fixed order:
// deterministic run to run)
template <int N>
__global__ void __launch_bounds__(256) small_scan_tau(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    __shared__ double s_buf[32 * 4];
    __shared__ int s_last;
    const int K = p.K, G = p.G, NT = p.NT;
    const cplx* Ql = a.Pseg + (size_t)(a.NSEG - 1) * NN * G;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) {
        const int g = p.gen[k];
        cplx x[N], y[N];
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = p.psi0[(size_t)i * K + k];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j) cfma(acc, __ldg(&Ql[(size_t)(i * N + j) * G + g]), x[j]);
            y[i] = acc;
            p.psi[((size_t)NT * N + i) * K + k] = acc;             // fw_propagators[k].state after the sweep
        }
        cplx t = mk(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < N; ++i) cfmac(t, p.tgt[(size_t)i * K + k], y[i]);
        p.tau[k] = t;
        p.jb[k] = 0.0;
        const double w = p.w ? p.w[k] : 1.0;
        v[0] = w * t.x;
        v[1] = w * t.y;
        v[2] = w * cnorm2(t);
    }
    block_sum<4>(v, s_buf);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) a.tau_part[(size_t)blockIdx.x * 4 + q] = v[q];
        __threadfence();
        s_last = atomicAdd(a.tau_ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 4) {
        __threadfence();
        double s = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(&a.tau_part[(size_t)b * 4 + threadIdx.x]);
        p.sums[threadIdx.x] = threadIdx.x == 3 ? 0.0 : s;
        if (threadIdx.x == 0) *a.tau_ticket = 0;
    }
}

// thread per (k, seg), k fastest: everything the two boundary chains produced.  fwd_only: fw_storage boundaries only
// (lazy read-back after a functional-only call).
template <int N>
__global__ void __launch_bounds__(128) small_scan_bounds(DevP p, SegArgs a, const cplx* __restrict__ chi_host, int fwd_only) {
    constexpr int NN = N * N;
    const int K = p.K, G = p.G, NT = p.NT;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)K * a.NSEG) return;
    const int k = (int)(idx % K), seg = (int)(idx / K);
    const int g = p.gen[k];
    const int nb = min(NT, (seg + 1) * a.S);
    cplx Q[NN];
#pragma unroll
    for (int c = 0; c < NN; ++c) Q[c] = __ldg(&a.Pseg[((size_t)seg * NN + c) * G + g]);
    cplx x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = p.psi0[(size_t)i * K + k];
    if (seg == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) st_cs(&p.psi[(size_t)i * K + k], x[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        cplx acc = mk(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < N; ++j) cfma(acc, Q[i * N + j], x[j]);
        st_cs(&p.psi[((size_t)nb * N + i) * K + k], acc);
    }
    if (fwd_only) return;
    // chi_k(T): optimize.jl:845-855 with the analytic chi of J_T_sm / J_T_re / J_T_ss, or the host's chi
    if (chi_host) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = chi_host[(size_t)k * N + i];
    } else {
        const double w = p.w ? p.w[k] : 1.0;
        const double Kg = (double)p.Kglobal;
        cplx c;
        if (p.functional == 0) c = mk(w * p.sums[0] / (Kg * Kg), w * p.sums[1] / (Kg * Kg));
        else if (p.functional == 1) c = mk(w / (2.0 * Kg), 0.0);
        else { cplx t = p.tau[k]; c = mk(w * t.x / Kg, w * t.y / Kg); }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = cmul(c, p.tgt[(size_t)i * K + k]);
    }
    double rho = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) rho += cnorm2(x[i]);
    rho = sqrt(rho);
    if (!(rho >= p.chi_min_norm)) {   // optimize.jl:1021-1025
        if (seg == 0 && atomicCAS(&p.flags->chi_bad_k, 0, k + 1) == 0) p.flags->chi_bad_rho = rho;
        rho = 1.0;
    }
    const double ir = 1.0 / rho;
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = cscale(x[i], ir);
    if (seg == 0) {
        p.rho[k] = rho;
#pragma unroll
        for (int i = 0; i < N; ++i) p.chiT[(size_t)k * N + i] = x[i];
    }
    if (seg < a.NSEG - 1) {   // chi(end of seg) = Q_seg Q_last^dagger chi(T); the last segment ends at T itself
        cplx y[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j)
                cfmac(acc, __ldg(&a.Pseg[((size_t)(a.NSEG - 1) * NN + j * N + i) * G + g]), x[j]);   // (Q_last^dagger x)_i
            y[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j) cfma(acc, Q[i * N + j], y[j]);
            x[i] = acc;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) a.chiE[((size_t)seg * N + i) * K + k] = x[i];
}
