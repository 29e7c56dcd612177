// small_seg.cuh -- time-segmented ("scan") schedule of the N <= 4 path.
//
// The time axis is strictly sequential (reference src/optimize.jl:731, 880), but a
// chain of NT dependent 3x3 mat-vecs per trajectory is pure latency on a GPU.  The
// propagators U_n only depend on eps_n, so the chain is cut into NSEG segments of S
// steps:
//   A1 small_form_U      (small_n.cuh)  all U_{g,n} in parallel
//   A2 small_segprod     P_{g,seg} = U_{n1-1} ... U_{n0}          parallel over (g, seg)
//   B1 small_segchain_fwd   Psi_k at the segment boundaries, tau_k  chain of NSEG steps
//   B2 small_segchain_bwd   chi_k(T) boundary condition (optimize.jl:845-869) and chi_k at
//                           the segment ends                        chain of NSEG steps
//   C1 small_segfwd      fills fw_storage inside every segment      parallel over (k, seg)
//   C2 small_seggrad     per (k, seg): walks the segment backwards carrying chi in
//                        registers; per step the GradGenerator contraction
//                        tau_grad[k][n,l] = rho_k <chi'_l | Psi_k(t_{n-1})>  (optimize.jl:893-895)
//                        and chi <- exp(+i H^dagger dt) chi come from the same Krylov vectors.
//
// Contraction used by C2 when the step needs at most SEG_MMAX Taylor orders and no
// sub-stepping (the usual piecewise-constant-control regime ||H dt|| << 1):
//   <chi'_l|Psi> = Tr(E_l^dagger M),  M = sum_{a,b} beta(a,b) bh_a ch_b^dagger,
//   bh_a = (-i H dt)^a Psi / a!,  ch_b = (+i H^dagger dt)^b chi / b!,  beta(a,b) = a! b!/(a+b+1)!
// i.e. the Frechet derivative of the exponential (docs/src/background.md:447-494) written with
// two Krylov sequences instead of the (1+2L) mat-vecs per order of the block recursion; the cost
// is independent of the number of controls up to the final traces.  Otherwise C2 falls back to
// the block recursion of small_gradient (small_n.cuh), which also serves gradient_method=:taylor.
#pragma once
#include "common.cuh"
#include "small_n.cuh"

constexpr int SEG_MMAX = 8;

// beta(a,b) = a! b! / (a+b+1)!
struct BetaTable {
    double v[SEG_MMAX][SEG_MMAX];
    constexpr BetaTable() : v() {
        for (int a = 0; a < SEG_MMAX; ++a)
            for (int b = 0; b < SEG_MMAX; ++b) {
                // a! b!/(a+b+1)! = 1/((a+b+1) * C(a+b, a))
                double c = 1.0;
                for (int t = 1; t <= a; ++t) c = c * (double)(b + t) / (double)t;
                v[a][b] = 1.0 / ((double)(a + b + 1) * c);
            }
    }
};
__constant__ BetaTable c_beta = BetaTable();

// gamma(a,b) = 1/(a+b+1): coefficients of the same contraction written with the Krylov vectors of the state at the
// END of the step (Hermitian generators, see seg_step_krylov_h)
struct GammaTable {
    double v[SEG_MMAX + 1][SEG_MMAX];
    constexpr GammaTable() : v() {
        for (int a = 0; a <= SEG_MMAX; ++a)
            for (int b = 0; b < SEG_MMAX; ++b) v[a][b] = 1.0 / (double)(a + b + 1);
    }
};
__constant__ GammaTable c_gamma = GammaTable();

struct SegArgs {
    int S, NSEG;
    int store_U;  // formseg keeps the propagators of every step in HBM (needed by the segment fill C1)
    int herm;     // all generators Hermitian: C2 recomputes the forward states backwards, C1 is not run
    cplx* Pseg;   // [NSEG][NN][G]
    cplx* chiE;   // [NSEG][N][K]   chi_k at the END time point of each segment
    int BKL;      // lanes of a warp that enumerate trajectories (power of two <= 32)
    // real-symmetric generators (small_sym.cuh): real copies of the operators and the device-side eligibility flag
    const double* H0r;   // [NN][G]
    const double* Hcr;   // [L][NN][G]
    int* notfast;        // [1] set by small_formseg_sym when a step needs > SEG_MMAX orders or sub-stepping
    // scan schedule (small_sym.cuh): Pseg holds the PREFIX products Q_seg = P_seg .. P_0
    int scan;            // the gradient kernel derives its boundary states from Q itself (no chains, no bounds kernel)
    const cplx* chi_host;// host chi of the current backward call (nullptr: analytic chi of the built-in functional)
    double* tau_part;    // [blocks][4] partial sums of small_scan_tau
    int* tau_ticket;     // [1] arrival counter of small_scan_tau (left at 0)
};

// ---------------------------------------------------------------------------
// A2: segment propagators
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) small_segprod(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int G = p.G, NT = p.NT;
    if (idx >= (long long)G * a.NSEG) return;
    const int g = (int)(idx % G), seg = (int)(idx / G);
    const int n0 = seg * a.S, n1 = min(NT, n0 + a.S);
    cplx P[NN], Un[NN];
    const cplx* Ug = p.U + g;
#pragma unroll
    for (int c = 0; c < NN; ++c) P[c] = ld_cs(&Ug[((size_t)n0 * NN + c) * G]);
    if (n0 + 1 < n1) {
#pragma unroll
        for (int c = 0; c < NN; ++c) Un[c] = ld_cs(&Ug[((size_t)(n0 + 1) * NN + c) * G]);
    }
    for (int n = n0 + 1; n < n1; ++n) {
        cplx Uc[NN];
#pragma unroll
        for (int c = 0; c < NN; ++c) Uc[c] = Un[c];
        if (n + 1 < n1) {
#pragma unroll
            for (int c = 0; c < NN; ++c) Un[c] = ld_cs(&Ug[((size_t)(n + 1) * NN + c) * G]);
        }
        cplx T[NN];
        sm_mm<N>(T, Uc, P);
#pragma unroll
        for (int c = 0; c < NN; ++c) P[c] = T[c];
    }
    cplx* o = a.Pseg + (size_t)seg * NN * G + g;
#pragma unroll
    for (int c = 0; c < NN; ++c) o[(size_t)c * G] = P[c];
}

// A1+A2 fused: thread per (g, seg) forms the S propagators of its segment one after the
// other (stored for C1) and accumulates their product, so U is not re-read from HBM.
// Used when G * NSEG alone fills the GPU; otherwise A1 (parallel over all (g, n)) + A2.
template <int N>
__global__ void __launch_bounds__(128) small_formseg(DevP p, SegArgs a) {
    constexpr int NN = N * N, BS = SmallCfg<N>::BS;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int G = p.G, NT = p.NT;
    if (idx >= (long long)G * a.NSEG) return;
    const int g = (int)(idx % G), seg = (int)(idx / G);
    const int n0 = seg * a.S, n1 = min(NT, n0 + a.S);
    cplx Pacc[NN];
    for (int n = n0; n < n1; ++n) {
        const double dt = p.tlist[n + 1] - p.tlist[n];
        cplx P[BS][NN];
        cplx X[NN];
#pragma unroll
        for (int c = 0; c < NN; ++c) X[c] = __ldg(&p.H0[(size_t)c * G + g]);
        for (int l = 0; l < p.L; ++l) {
            double am = p.eps[l * NT + n];
            if (p.shape) am *= p.shape[l * NT + n];
#pragma unroll
            for (int c = 0; c < NN; ++c) cfmar(X[c], am, __ldg(&p.Hc[((size_t)l * NN + c) * G + g]));
        }
#pragma unroll
        for (int c = 0; c < NN; ++c) P[0][c] = mk(dt * X[c].y, -dt * X[c].x);   // -i dt H
        int degree, s;
        exp_plan(sm_norm1<N>(P[0]), degree, s);
        if (s > 0) {
            const double sc = ldexp(1.0, -s);
#pragma unroll
            for (int c = 0; c < NN; ++c) P[0][c] = cscale(P[0][c], sc);
        }
        sm_expm<N, BS>(P, X, degree, s);
        if (a.store_U) {
            cplx* Uo = p.U + (size_t)n * NN * G + g;
#pragma unroll
            for (int c = 0; c < NN; ++c) st_cs(&Uo[(size_t)c * G], X[c]);
        }
        if (n == n0) {
#pragma unroll
            for (int c = 0; c < NN; ++c) Pacc[c] = X[c];
        } else {
            sm_mm<N>(P[0], X, Pacc);
#pragma unroll
            for (int c = 0; c < NN; ++c) Pacc[c] = P[0][c];
        }
    }
    cplx* o = a.Pseg + (size_t)seg * NN * G + g;
#pragma unroll
    for (int c = 0; c < NN; ++c) o[(size_t)c * G] = Pacc[c];
}

// ---------------------------------------------------------------------------
// B1: forward chain over segment boundaries
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(64) small_segchain_fwd(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    const int K = p.K, G = p.G, NT = p.NT;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int g = p.gen[k];
    const cplx* Pg = a.Pseg + g;
    cplx psi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        psi[i] = p.psi0[(size_t)i * K + k];
        st_cs(&p.psi[(size_t)i * K + k], psi[i]);
    }
    // The segment propagators do not depend on the state: the loads of CH segments are issued together (one L2
    // round trip per CH steps of the chain instead of one per step), then applied one after the other.
    constexpr int CH = N <= 3 ? 4 : 2;
    for (int seg0 = 0; seg0 < a.NSEG; seg0 += CH) {
        cplx Pc[CH][NN];
#pragma unroll
        for (int u = 0; u < CH; ++u)
            if (seg0 + u < a.NSEG) {
#pragma unroll
                for (int c = 0; c < NN; ++c) Pc[u][c] = __ldg(&Pg[((size_t)(seg0 + u) * NN + c) * G]);
            }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int seg = seg0 + u;
            if (seg < a.NSEG) {
                cplx nw[N];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    cplx acc = mk(0.0, 0.0);
#pragma unroll
                    for (int j = 0; j < N; ++j) cfma(acc, Pc[u][i * N + j], psi[j]);
                    nw[i] = acc;
                }
                const int nb = min(NT, (seg + 1) * a.S);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    psi[i] = nw[i];
                    st_cs(&p.psi[((size_t)nb * N + i) * K + k], psi[i]);
                }
            }
        }
    }
    cplx acc = mk(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < N; ++i) cfmac(acc, p.tgt[(size_t)i * K + k], psi[i]);
    p.tau[k] = acc;
    p.jb[k] = 0.0;
}

// ---------------------------------------------------------------------------
// B2: chi boundary condition + backward chain over segment boundaries
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(64) small_segchain_bwd(DevP p, SegArgs a, const cplx* __restrict__ chi_host) {
    constexpr int NN = N * N;
    const int K = p.K, G = p.G;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int g = p.gen[k];
    const cplx* Pg = a.Pseg + g;
    cplx x[N];
    if (chi_host) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = chi_host[(size_t)k * N + i];
    } else {   // optimize.jl:845-855 with the analytic chi of J_T_sm / J_T_re / J_T_ss
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
        if (atomicCAS(&p.flags->chi_bad_k, 0, k + 1) == 0) p.flags->chi_bad_rho = rho;
        rho = 1.0;
    }
    const double ir = 1.0 / rho;
    p.rho[k] = rho;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        x[i] = cscale(x[i], ir);
        p.chiT[(size_t)k * N + i] = x[i];
    }
    // chi at the END of segment `seg` is stored, then chi <- P_seg^dagger chi; P of segment 0 is never needed.
    // As in the forward chain the propagators of CH segments are loaded together.
    constexpr int CH = N <= 3 ? 4 : 2;
    for (int seg0 = a.NSEG - 1; seg0 >= 0; seg0 -= CH) {
        cplx Pc[CH][NN];
#pragma unroll
        for (int u = 0; u < CH; ++u)
            if (seg0 - u > 0) {
#pragma unroll
                for (int c = 0; c < NN; ++c) Pc[u][c] = __ldg(&Pg[((size_t)(seg0 - u) * NN + c) * G]);
            }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int seg = seg0 - u;
            if (seg >= 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) a.chiE[((size_t)seg * N + i) * K + k] = x[i];
                if (seg > 0) {
                    cplx nw[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        cplx acc = mk(0.0, 0.0);
#pragma unroll
                        for (int j = 0; j < N; ++j) cfmac(acc, Pc[u][j * N + i], x[j]);   // (P^dagger x)_i
                        nw[i] = acc;
                    }
#pragma unroll
                    for (int i = 0; i < N; ++i) x[i] = nw[i];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// B1 + B2 in one pass (gradient calls with a built-in functional): chi_k(T) = c_k tgt_k with a scalar c_k
// (optimize.jl:845-855), and the backward chain is linear, so tgt_k is carried backwards over the segment
// boundaries WHILE Psi_k is carried forwards -- two independent dependency chains per thread, the latency of one.
// chiE receives the propagated RAW targets; the gradient kernel multiplies by c_k / rho_k (SegArgs::scan == 2).
// ---------------------------------------------------------------------------
// The segment propagators do not depend on the states: each thread streams ITS propagators (forward and backward
// order) through a private column of a cp.async ring in shared memory, CHAIN_D - 1 segments ahead, so that a chain
// step costs one dependent mat-vec instead of one L2 round trip (measured: 25 us -> see profiles/ for 45 segments).
constexpr int CHAIN_D = 8;     // ring depth
constexpr int CHAIN_BD = 32;   // trajectories per block; the block has 2 warps: warp 0 = forward chain, warp 1 = backward chain
inline size_t chain_ring_bytes(int N) { return (size_t)2 * CHAIN_D * N * N * CHAIN_BD * sizeof(cplx); }

// With one warp per SM the chain is bound by the issue latency of its own instruction stream (ncu: 0.26 IPC, 244
// instructions per segment for both directions), so the two directions get a warp each: same 32 trajectories, half
// the instructions per warp, no synchronisation between them.
template <int N>
__global__ void __launch_bounds__(2 * CHAIN_BD) small_segchain_dual(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    extern __shared__ __align__(16) cplx ring_all[];   // [2 roles][CHAIN_D][NN][CHAIN_BD]
    const int K = p.K, G = p.G, NT = p.NT;
    const int lane = threadIdx.x & (CHAIN_BD - 1);
    const bool bwd = threadIdx.x >= CHAIN_BD;
    cplx* ring = ring_all + (bwd ? (size_t)CHAIN_D * NN * CHAIN_BD : 0) + lane;
    const int k = blockIdx.x * CHAIN_BD + lane;
    const int kk = k < K ? k : K - 1;              // idle lanes stream valid addresses and store nothing
    const bool live = k < K;
    const int g = p.gen[kk];
    const cplx* Pg = a.Pseg + g;
    const size_t segstride = (size_t)NN * G;
    // running source pointer (one 64-bit add per element): forward role walks the segments upwards, backward role downwards
    const cplx* src = bwd ? Pg + (size_t)(a.NSEG - 1) * segstride : Pg;
    const int nuse = bwd ? a.NSEG - 1 : a.NSEG;    // the backward role never applies P of segment 0
    auto issue = [&](int s) {
        if (s < nuse) {
            cplx* st = ring + (size_t)(s % CHAIN_D) * NN * CHAIN_BD;
            const cplx* q = src;
#pragma unroll
            for (int c = 0; c < NN; ++c) { cp_async16(st + c * CHAIN_BD, q); q += G; }
            if (bwd) src -= segstride; else src += segstride;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < CHAIN_D - 1; ++s) issue(s);
    cplx x[N];
    if (!bwd) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            x[i] = p.psi0[(size_t)i * K + kk];
            if (live) st_cs(&p.psi[(size_t)i * K + k], x[i]);
        }
        for (int sf = 0; sf < a.NSEG; ++sf) {
            cp_async_wait<CHAIN_D - 2>();          // this thread's copies of segment sf have landed
            const cplx* st = ring + (size_t)(sf % CHAIN_D) * NN * CHAIN_BD;
            cplx nw[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < N; ++j) cfma(acc, st[(i * N + j) * CHAIN_BD], x[j]);
                nw[i] = acc;
            }
            const int nb = min(NT, (sf + 1) * a.S);
            cplx* o = p.psi + (size_t)nb * N * K + k;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                x[i] = nw[i];
                if (live) st_cs(o, x[i]);
                o += K;
            }
            issue(sf + CHAIN_D - 1);               // refill the slot the previous segment used
        }
        if (live) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < N; ++i) cfmac(acc, p.tgt[(size_t)i * K + k], x[i]);
            p.tau[k] = acc;
            p.jb[k] = 0.0;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = p.tgt[(size_t)i * K + kk];
        for (int it = 0; it < a.NSEG; ++it) {
            const int sb = a.NSEG - 1 - it;
            // the target at the END of segment sb, then through P_sb^dagger
            if (live) {
                cplx* o = a.chiE + (size_t)sb * N * K + k;
#pragma unroll
                for (int i = 0; i < N; ++i) { *o = x[i]; o += K; }
            }
            if (sb > 0) {
                cp_async_wait<CHAIN_D - 2>();
                const cplx* st = ring + (size_t)(it % CHAIN_D) * NN * CHAIN_BD;
                cplx ny[N];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    cplx acc = mk(0.0, 0.0);
#pragma unroll
                    for (int j = 0; j < N; ++j) cfmac(acc, st[(j * N + i) * CHAIN_BD], x[j]);
                    ny[i] = acc;
                }
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = ny[i];
                issue(it + CHAIN_D - 1);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// C1: forward states inside every segment (thread per (k, seg), k fastest)
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) small_segfwd(DevP p, SegArgs a) {
    constexpr int NN = N * N;
    const int K = p.K, G = p.G, NT = p.NT;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)K * a.NSEG) return;
    const int k = (int)(idx % K), seg = (int)(idx / K);
    const int n0 = seg * a.S, n1 = min(NT, n0 + a.S);
    if (n0 + 1 >= n1) return;   // single-step segment: both ends come from the chain
    const int g = p.gen[k];
    const cplx* Ug = p.U + g;
    cplx psi[N], Un[NN];
#pragma unroll
    for (int c = 0; c < NN; ++c) Un[c] = ld_cs(&Ug[((size_t)n0 * NN + c) * G]);
#pragma unroll
    for (int i = 0; i < N; ++i) psi[i] = p.psi[((size_t)n0 * N + i) * K + k];
    for (int n = n0; n < n1 - 1; ++n) {   // the state at n1 was written by the chain
        cplx Uc[NN];
#pragma unroll
        for (int c = 0; c < NN; ++c) Uc[c] = Un[c];
        if (n + 1 < n1 - 1) {
#pragma unroll
            for (int c = 0; c < NN; ++c) Un[c] = ld_cs(&Ug[((size_t)(n + 1) * NN + c) * G]);
        }
        cplx nw[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j) cfma(acc, Uc[i * N + j], psi[j]);
            nw[i] = acc;
        }
        cplx* o = p.psi + ((size_t)(n + 1) * N) * K + k;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            psi[i] = nw[i];
            st_cs(o + (size_t)i * K, psi[i]);
        }
    }
}

// ---------------------------------------------------------------------------
// C2: backward walk through a segment fused with the gradient contraction
// ---------------------------------------------------------------------------
// one step, Krylov form. A = +i dt H^dagger (row-major). On exit chi <- exp(A) chi and
// M = sum beta(a,b) bh_a ch_b^dagger.
template <int N>
GB_D void seg_step_krylov(const cplx (&A)[N * N], const cplx (&psi)[N], cplx (&chi)[N], cplx (&M)[N * N], int m) {
    cplx e[SEG_MMAX][N];
    cplx bv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) bv[i] = psi[i];
#pragma unroll
    for (int b = 0; b < SEG_MMAX; ++b)
#pragma unroll
        for (int i = 0; i < N; ++i) e[b][i] = cscale(bv[i], c_beta.v[0][b]);
#pragma unroll
    for (int aa = 1; aa < SEG_MMAX; ++aa) {
        if (aa < m) {
            // bv <- (A^dagger bv)/aa   (A^dagger = -i dt H)
            cplx nb[N];
            const double inv = 1.0 / (double)aa;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < N; ++q) cfmac(acc, A[q * N + i], bv[q]);
                nb[i] = cscale(acc, inv);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) bv[i] = nb[i];
#pragma unroll
            for (int b = 0; b < SEG_MMAX - aa; ++b)
#pragma unroll
                for (int i = 0; i < N; ++i) cfmar(e[b][i], c_beta.v[aa][b], bv[i]);
        }
    }
    cplx cv[N], acc_chi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { cv[i] = chi[i]; acc_chi[i] = chi[i]; }
#pragma unroll
    for (int c = 0; c < N * N; ++c) M[c] = mk(0.0, 0.0);
#pragma unroll
    for (int b = 0; b < SEG_MMAX; ++b) {
        if (b < m) {
#pragma unroll
            for (int pp = 0; pp < N; ++pp)
#pragma unroll
                for (int q = 0; q < N; ++q) {   // M_pq += e_b[p] * conj(cv[q])
                    cplx& t = M[pp * N + q];
                    const cplx x = e[b][pp], y = cv[q];
                    t.x = fma(x.x, y.x, t.x); t.x = fma(x.y, y.y, t.x);
                    t.y = fma(x.y, y.x, t.y); t.y = fma(-x.x, y.y, t.y);
                }
            cplx nc[N];
            const double inv = 1.0 / (double)(b + 1);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < N; ++q) cfma(acc, A[i * N + q], cv[q]);
                nc[i] = cscale(acc, inv);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) { cv[i] = nc[i]; acc_chi[i] = cadd(acc_chi[i], nc[i]); }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) chi[i] = acc_chi[i];
}

// Hermitian generators: the forward state is carried backwards through the segment next to chi instead of being
// read from fw_storage (H = H^dagger, so exp(-iH dt)^{-1} = exp(+iH dt) = exp(A) with the same A = +i dt H^dagger):
//   bt_a = A^a Psi(t_n)/a!,  ch_b = A^b chi(t_n)/b!,  Psi(t_{n-1}) = sum_a bt_a,  chi(t_{n-1}) = sum_b ch_b,
//   M = int_0^1 Psi(s) chi(s)^dagger ds = sum_{a,b} bt_a ch_b^dagger / (a+b+1)
// (same M as seg_step_krylov, expanded around the end of the step). On exit psi and chi hold the start-of-step states.
template <int N>
GB_D void seg_step_krylov_h(const cplx (&A)[N * N], cplx (&psi)[N], cplx (&chi)[N], cplx (&M)[N * N], int m) {
    cplx e[SEG_MMAX][N];
    cplx bv[N], acc_psi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { bv[i] = psi[i]; acc_psi[i] = psi[i]; }
#pragma unroll
    for (int b = 0; b < SEG_MMAX; ++b)
#pragma unroll
        for (int i = 0; i < N; ++i) e[b][i] = cscale(bv[i], c_gamma.v[0][b]);
#pragma unroll
    for (int aa = 1; aa <= SEG_MMAX; ++aa) {
        if (aa <= m) {
            cplx nb[N];
            const double inv = 1.0 / (double)aa;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < N; ++q) cfma(acc, A[i * N + q], bv[q]);
                nb[i] = cscale(acc, inv);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) { bv[i] = nb[i]; acc_psi[i] = cadd(acc_psi[i], nb[i]); }
            if (aa < m) {
#pragma unroll
                for (int b = 0; b < SEG_MMAX - aa; ++b)
#pragma unroll
                    for (int i = 0; i < N; ++i) cfmar(e[b][i], c_gamma.v[aa][b], bv[i]);
            }
        }
    }
    cplx cv[N], acc_chi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { cv[i] = chi[i]; acc_chi[i] = chi[i]; }
#pragma unroll
    for (int c = 0; c < N * N; ++c) M[c] = mk(0.0, 0.0);
#pragma unroll
    for (int b = 0; b < SEG_MMAX; ++b) {
        if (b < m) {
#pragma unroll
            for (int pp = 0; pp < N; ++pp)
#pragma unroll
                for (int q = 0; q < N; ++q) {   // M_pq += e_b[p] * conj(cv[q])
                    cplx& t = M[pp * N + q];
                    const cplx x = e[b][pp], y = cv[q];
                    t.x = fma(x.x, y.x, t.x); t.x = fma(x.y, y.y, t.x);
                    t.y = fma(x.y, y.x, t.y); t.y = fma(-x.x, y.y, t.y);
                }
            cplx nc[N];
            const double inv = 1.0 / (double)(b + 1);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < N; ++q) cfma(acc, A[i * N + q], cv[q]);
                nc[i] = cscale(acc, inv);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) { cv[i] = nc[i]; acc_chi[i] = cadd(acc_chi[i], nc[i]); }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) { chi[i] = acc_chi[i]; psi[i] = acc_psi[i]; }
}

// v <- exp(f A)^nsub v by the m-term Taylor recursion per sub-step
template <int N>
GB_D void seg_apply_exp(const cplx (&A)[N * N], cplx (&v)[N], int m, int nsub, double f) {
    for (int sub = 0; sub < nsub; ++sub) {
        cplx ta[N];
#pragma unroll
        for (int i = 0; i < N; ++i) ta[i] = v[i];
        for (int j = 1; j <= m; ++j) {
            const double inv = f / (double)j;
            cplx na[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < N; ++q) cfma(acc, A[i * N + q], ta[q]);
                na[i] = cscale(acc, inv);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) { ta[i] = na[i]; v[i] = cadd(v[i], na[i]); }
        }
    }
}

template <int N, int LC, bool HERM>
__global__ void __launch_bounds__(128) small_seggrad(DevP p, SegArgs a, int l0, const int* __restrict__ run_if) {
    constexpr int NN = N * N;
    if (run_if && !*run_if) return;   // uniform over the grid: small_seggrad_sym served this call
    const int K = p.K, G = p.G, NT = p.NT;
    const int lane = threadIdx.x & 31;
    const int BKL = a.BKL, SPW = 32 / BKL;                 // segments per warp
    const int KGR = (K + BKL - 1) / BKL;                   // trajectory groups
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int SGR = (a.NSEG + SPW - 1) / SPW;              // segment groups
    if (wid >= (long long)KGR * SGR) return;               // whole warp exits
    const int kg = (int)(wid % KGR), sg = (int)(wid / KGR);
    const int tk = lane % BKL, ts = lane / BKL;
    const int k = kg * BKL + tk, seg = sg * SPW + ts;
    const bool live = (k < K) && (seg < a.NSEG);
    const int kk = k < K ? k : K - 1;
    const int sseg = seg < a.NSEG ? seg : a.NSEG - 1;
    const int n0 = sseg * a.S, n1 = min(NT, n0 + a.S);
    const int g = p.gen[kk];
    const double rho = p.rho[kk];
    const bool taylor = p.grad_method != 0;

    cplx chi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) chi[i] = a.chiE[((size_t)sseg * N + i) * K + kk];
    // general generators: psin prefetches Psi(t_{n-1}) of the next step from fw_storage;
    // Hermitian generators: psin carries the state at the END of the current step, starting from the segment
    // boundary written by the forward chain, and is propagated backwards together with chi
    cplx psin[N];
#pragma unroll
    for (int i = 0; i < N; ++i) psin[i] = ld_cs(&p.psi[((size_t)(HERM ? n1 : n1 - 1) * N + i) * K + kk]);

    for (int st = 0; st < a.S; ++st) {                     // uniform trip count across the warp
        const int n = n1 - 1 - st;
        const bool act = live && n >= n0;
        const int nn = n >= n0 ? n : n0;
        cplx psi[N];
#pragma unroll
        for (int i = 0; i < N; ++i) psi[i] = psin[i];
        if (!HERM && nn - 1 >= n0) {
#pragma unroll
            for (int i = 0; i < N; ++i) psin[i] = ld_cs(&p.psi[((size_t)(nn - 1) * N + i) * K + kk]);
        }
        const double dt = p.tlist[nn + 1] - p.tlist[nn];
        // A = H^dagger (then scaled to +i dt H^dagger)
        cplx A[NN];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) A[i * N + j] = __ldg(&p.H0[(size_t)(j * N + i) * G + g]);
        for (int l = 0; l < p.L; ++l) {
            double am = p.eps[l * NT + nn];
            if (p.shape) am *= p.shape[l * NT + nn];
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j)
                    cfmar(A[i * N + j], am, __ldg(&p.Hc[((size_t)l * NN + j * N + i) * G + g]));
        }
        int m, s;
        {
            const double nrm = dt * sm_norm1<N>(A);
            if (!taylor) vec_plan(nrm, m, s);
            else { s = 0; m = p.taylor_max_order; }
        }
        const double sc = dt * ldexp(1.0, -s);
#pragma unroll
        for (int c = 0; c < NN; ++c) A[c] = mk(sc * A[c].y, sc * A[c].x);   // i*sc*conj(H_ji): A held H transposed
        double red[LC];
        cplx tg[LC];
        // N = 4 does not fit the Krylov form in registers: block recursion only
        const bool fast = (N <= 3) && __all_sync(0xffffffffu, !taylor && s == 0 && m <= SEG_MMAX);
        if (N <= 3 && fast) {
            cplx M[NN];
            if (HERM) {
                seg_step_krylov_h<N>(A, psi, chi, M, m);
#pragma unroll
                for (int i = 0; i < N; ++i) psin[i] = psi[i];
            } else {
                seg_step_krylov<N>(A, psi, chi, M, m);
            }
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                double sl = sc;
                if (p.dshape) sl *= p.dshape[(l0 + l) * NT + nn];
                // E_pq = i*sl*conj(Hc[q][p]);  Tr(E^dagger M) = sum conj(E_pq) M_pq
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int pp = 0; pp < N; ++pp)
#pragma unroll
                    for (int q = 0; q < N; ++q) {
                        const cplx h = __ldg(&p.Hc[((size_t)(l0 + l) * NN + q * N + pp) * G + g]);
                        cfmac(acc, mk(sl * h.y, sl * h.x), M[pp * N + q]);
                    }
                tg[l] = cscale(acc, rho);
            }
        } else {
            if (HERM) {   // Psi(t_{n-1}) = exp(+i H dt) Psi(t_n), then the block recursion as for general generators
                if (!taylor) seg_apply_exp<N>(A, psi, m, 1 << s, 1.0);
                else {
                    int mm, ss;
                    vec_plan(sm_norm1<N>(A) * (dt / sc), mm, ss);
                    seg_apply_exp<N>(A, psi, mm, 1 << ss, ldexp(1.0, -ss));
                }
#pragma unroll
                for (int i = 0; i < N; ++i) psin[i] = psi[i];
            }
            // block (GradGenerator) recursion, identical to small_gradient (small_n.cuh)
            cplx E[LC][NN];
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                double sl = sc;
                if (p.dshape) sl *= p.dshape[(l0 + l) * NT + nn];
#pragma unroll
                for (int i = 0; i < N; ++i)
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const cplx h = __ldg(&p.Hc[((size_t)(l0 + l) * NN + j * N + i) * G + g]);
                        E[l][i * N + j] = mk(sl * h.y, sl * h.x);
                    }
            }
            cplx asum[N], bsum[LC][N];
#pragma unroll
            for (int i = 0; i < N; ++i) asum[i] = chi[i];
#pragma unroll
            for (int l = 0; l < LC; ++l)
#pragma unroll
                for (int i = 0; i < N; ++i) bsum[l][i] = mk(0.0, 0.0);
            bool converged = !taylor || !p.taylor_check;
            double rlast = 0.0;
            bool done[LC];   // per-control early return of taylor_grad_step! (optimize.jl:633-638)
#pragma unroll
            for (int l = 0; l < LC; ++l) done[l] = false;
            for (int sub = 0; sub < (1 << s); ++sub) {
                cplx ta[N], tb[LC][N];
#pragma unroll
                for (int i = 0; i < N; ++i) ta[i] = asum[i];
#pragma unroll
                for (int l = 0; l < LC; ++l)
#pragma unroll
                    for (int i = 0; i < N; ++i) tb[l][i] = bsum[l][i];
                for (int j = 1; j <= m; ++j) {
                    const double inv = 1.0 / (double)j;
#pragma unroll
                    for (int l = 0; l < LC; ++l) {
                        cplx nb[N];
                        double r2 = 0.0;
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            cplx acc = mk(0.0, 0.0);
#pragma unroll
                            for (int q = 0; q < N; ++q) cfma(acc, E[l][i * N + q], ta[q]);
#pragma unroll
                            for (int q = 0; q < N; ++q) cfma(acc, A[i * N + q], tb[l][q]);
                            nb[i] = cscale(acc, inv);
                            r2 += cnorm2(nb[i]);
                        }
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            tb[l][i] = nb[i];
                            if (!done[l]) bsum[l][i] = cadd(bsum[l][i], nb[i]);
                        }
                        if (taylor && p.taylor_check && j >= 2 && !done[l]) {
                            const double r = sqrt(r2);
                            rlast = r;
                            if (r < p.taylor_tol) done[l] = true;
                        }
                    }
                    cplx na[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        cplx acc = mk(0.0, 0.0);
#pragma unroll
                        for (int q = 0; q < N; ++q) cfma(acc, A[i * N + q], ta[q]);
                        na[i] = cscale(acc, inv);
                    }
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        ta[i] = na[i];
                        asum[i] = cadd(asum[i], na[i]);
                    }
                    if (taylor && p.taylor_check && j >= 2) {

                        bool all_done = true;
#pragma unroll

                        for (int l = 0; l < LC; ++l) all_done = all_done && done[l];

                        if (all_done) { converged = true; break; }

                    }
                }
            }
            if (act && !converged && p.taylor_max_order > 1) {   // optimize.jl:642-648
                if (atomicExch(&p.flags->taylor_fail, 1) == 0) p.flags->taylor_r = rlast;
            }
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int i = 0; i < N; ++i) cfmac(acc, bsum[l][i], psi[i]);
                tg[l] = cscale(acc, rho);
            }
            if (taylor) {
                // :taylor only defines chi'_l; chi itself is propagated with the full exponential
                // (prop_step!, optimize.jl:972): converged series independent of taylor_grad_tolerance
                int mm, ss;
                vec_plan(sm_norm1<N>(A) * (dt / sc), mm, ss);
                const double f = ldexp(1.0, -ss);
#pragma unroll
                for (int i = 0; i < N; ++i) asum[i] = chi[i];
                for (int sub = 0; sub < (1 << ss); ++sub) {
                    cplx ta[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) ta[i] = asum[i];
                    for (int j = 1; j <= mm; ++j) {
                        const double inv = f / (double)j;
                        cplx na[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            cplx acc = mk(0.0, 0.0);
#pragma unroll
                            for (int q = 0; q < N; ++q) cfma(acc, A[i * N + q], ta[q]);
                            na[i] = cscale(acc, inv);
                        }
#pragma unroll
                        for (int i = 0; i < N; ++i) { ta[i] = na[i]; asum[i] = cadd(asum[i], na[i]); }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) chi[i] = asum[i];
        }
#pragma unroll
        for (int l = 0; l < LC; ++l) {
            if (act && p.taugrads) p.taugrads[((size_t)k * p.L + (l0 + l)) * NT + n] = tg[l];
            red[l] = act ? tg[l].x : 0.0;
        }
        // fixed-order sum over the BKL trajectories of this lane group
#pragma unroll
        for (int l = 0; l < LC; ++l)
            for (int off = BKL >> 1; off > 0; off >>= 1)
                red[l] += __shfl_down_sync(0xffffffffu, red[l], off, BKL);
        if (tk == 0 && seg < a.NSEG && n >= n0) {
#pragma unroll
            for (int l = 0; l < LC; ++l)
                p.partial[(size_t)kg * p.L * NT + (size_t)(l0 + l) * NT + n] = red[l];
        }
    }
}
