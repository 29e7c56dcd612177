// warp_n.cuh -- 5 <= N <= 32 path: one sub-warp (W = 8, 16 or 32 lanes) per
// (generator, step) / trajectory / (trajectory, step); lane i owns row i of a
// state, matrices live in shared memory, vectors are exchanged with shuffles.
//
// Device layout (array-of-structures, a state / matrix is contiguous):
//   H0w [g][i*N + j]      Hcw [g][l][i*N + j]     Dw [d][i*N + j]
//   U   [n][g][i*N + j]
//   psi [n][k][i]  n = 0..NT        chi [n][k][i]  n = 1..NT
//
// Same phases as the small path (see small_n.cuh); propagator formation uses the
// Paterson-Stockmeyer Taylor evaluation with block size 2 on shared-memory matrices.
#pragma once
#include "common.cuh"
#include "../../include/grape_b200.h"
#include <algorithm>
#include <string>
#include <vector>

constexpr int WARP_MAX_N = 32;
constexpr int WARP_D = 4;   // cp.async ring depth of the chain kernels

struct WarpPlan {
    int W;            // sub-warp width
    int NS;           // padded row stride (odd) of ring-buffer matrices
    int spbA;         // sub-warps per block, phase A
    size_t smemA, smemB, smemC;
    int spbC;
};

template <int W>
GB_D unsigned sub_mask() {
    if constexpr (W == 32) {
        return 0xffffffffu;
    } else {
        const unsigned lane = threadIdx.x & 31;
        return ((1u << W) - 1u) << (lane & ~(W - 1));
    }
}
template <int W>
GB_D double sub_sum(double v, unsigned mask) {
#pragma unroll
    for (int off = W >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off, W);
    return v;
}
template <int W>
GB_D double sub_max(double v, unsigned mask) {
#pragma unroll
    for (int off = W >> 1; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, off, W));
    return v;
}
template <int W>
GB_D cplx sub_bcast(cplx v, int src, unsigned mask) {
    return mk(__shfl_sync(mask, v.x, src, W), __shfl_sync(mask, v.y, src, W));
}

// C = A*B (N x N, row-major, stride N) in shared memory; lane r owns column r.
// 4-row register tile: 3 shared loads per 2 complex FMA -> FP64 bound.
template <int W>
GB_D void sw_matmul(cplx* __restrict__ C, const cplx* __restrict__ A, const cplx* __restrict__ B,
                    int N, int r, unsigned mask) {
    if (r < N) {
        for (int i0 = 0; i0 < N; i0 += 4) {
            cplx acc[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t] = mk(0.0, 0.0);
            const int i1 = min(i0 + 1, N - 1), i2 = min(i0 + 2, N - 1), i3 = min(i0 + 3, N - 1);
            for (int k = 0; k < N; ++k) {
                const cplx b = B[k * N + r];
                cfma(acc[0], A[i0 * N + k], b);
                cfma(acc[1], A[i1 * N + k], b);
                cfma(acc[2], A[i2 * N + k], b);
                cfma(acc[3], A[i3 * N + k], b);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (i0 + t < N) C[(i0 + t) * N + r] = acc[t];
        }
    }
    __syncwarp(mask);
}

// ---------------------------------------------------------------------------
// Phase A
// ---------------------------------------------------------------------------
template <int W>
__global__ void warp_form_U(DevP p, int spb) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N = p.N, NN = N * N, G = p.G, NT = p.NT;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const long long unit = (long long)blockIdx.x * spb + sub;
    if (unit >= (long long)G * NT) return;   // whole sub-warp exits together
    const int g = (int)(unit % G), n = (int)(unit / G);
    cplx* A = reinterpret_cast<cplx*>(smraw) + (size_t)sub * 4 * NN;
    cplx* A2 = A + NN;
    cplx* X = A2 + NN;
    cplx* T = X + NN;
    const double dt = p.tlist[n + 1] - p.tlist[n];
    const cplx* H0 = p.H0 + (size_t)g * NN;
    const cplx* Hc = p.Hc + (size_t)g * p.L * NN;
    for (int e = r; e < NN; e += W) {
        cplx h = __ldg(&H0[e]);
        for (int l = 0; l < p.L; ++l) {
            double a = p.eps[l * NT + n];
            if (p.shape) a *= p.shape[l * NT + n];
            cfmar(h, a, __ldg(&Hc[(size_t)l * NN + e]));
        }
        A[e] = mk(dt * h.y, -dt * h.x);   // -i dt H
    }
    __syncwarp(mask);
    double cs = 0.0;
    if (r < N)
        for (int i = 0; i < N; ++i) cs += cabs1(A[i * N + r]);
    const double nrm = sub_max<W>(cs, mask);
    int degree, s;
    exp_plan(nrm, degree, s);
    if (s > 0) {
        const double sc = ldexp(1.0, -s);
        for (int e = r; e < NN; e += W) A[e] = cscale(A[e], sc);
        __syncwarp(mask);
    }
    sw_matmul<W>(A2, A, A, N, r, mask);
    const int q = (degree + 1) / 2 - 1;
    {
        const double c0 = c_invfact[2 * q], c1 = c_invfact[2 * q + 1];
        for (int e = r; e < NN; e += W) {
            cplx v = cscale(A[e], c1);
            if (e / N == e % N) v.x += c0;
            X[e] = v;
        }
        __syncwarp(mask);
    }
    for (int t = q - 1; t >= 0; --t) {
        sw_matmul<W>(T, A2, X, N, r, mask);
        const double c0 = c_invfact[2 * t], c1 = c_invfact[2 * t + 1];
        for (int e = r; e < NN; e += W) {
            cplx v = T[e];
            cfmar(v, c1, A[e]);
            if (e / N == e % N) v.x += c0;
            X[e] = v;
        }
        __syncwarp(mask);
    }
    for (int t = 0; t < s; ++t) {
        sw_matmul<W>(T, X, X, N, r, mask);
        cplx* tmp = X; X = T; T = tmp;
    }
    cplx* Uo = p.U + ((size_t)n * G + g) * NN;
    for (int e = r; e < NN; e += W) Uo[e] = X[e];
}

// (D x)_r for lane r, D row-major in global memory (L1/L2 resident)
template <int W>
GB_D cplx sw_rowdot_global(const cplx* __restrict__ Dm, int N, int r, cplx x, unsigned mask) {
    cplx acc = mk(0.0, 0.0);
    const int rr = r < N ? r : N - 1;
    for (int j = 0; j < N; ++j) {
        const cplx xj = sub_bcast<W>(x, j, mask);
        cfma(acc, __ldg(&Dm[(size_t)rr * N + j]), xj);
    }
    return acc;
}

// ---------------------------------------------------------------------------
// Phase B1 / B2: chains. One sub-warp per trajectory, U_n staged through a
// cp.async ring (row stride NS odd -> conflict-free row and column reads).
// ---------------------------------------------------------------------------
template <int W>
GB_D void sw_issue(cplx* __restrict__ dst, const cplx* __restrict__ src, int N, int NS, int r, bool valid) {
    if (valid) {
        const int NN = N * N;
        for (int e = r; e < NN; e += W) cp_async16(dst + (e / N) * NS + (e % N), src + e);
    }
    cp_async_commit();
}

template <int W>
__global__ void warp_forward(DevP p, int NS) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N = p.N, NN = N * N, G = p.G, NT = p.NT, K = p.K;
    const int spb = blockDim.x / W;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const int k = blockIdx.x * spb + sub;
    if (k >= K) return;
    const int g = p.gen[k];
    const size_t stage = (size_t)N * NS;
    cplx* ring = reinterpret_cast<cplx*>(smraw) + (size_t)sub * WARP_D * stage;
    const cplx* Ug = p.U + (size_t)g * NN;
    const size_t ustride = (size_t)G * NN;
#pragma unroll
    for (int n = 0; n < WARP_D - 1; ++n) sw_issue<W>(ring + (n % WARP_D) * stage, Ug + n * ustride, N, NS, r, n < NT);
    const bool own = r < N;
    cplx psi = own ? p.psi0[(size_t)k * N + r] : mk(0.0, 0.0);
    if (own) p.psi[(size_t)k * N + r] = psi;
    const bool gb = p.gb_kind != 0;
    const cplx* Dk = gb ? p.D + (p.gb_nD == 1 ? 0 : (size_t)k * NN) : nullptr;
    double jb = 0.0;
    if (gb) {
        const cplx t = sw_rowdot_global<W>(Dk, N, r, psi, mask);
        jb = sub_sum<W>(own ? psi.x * t.x + psi.y * t.y : 0.0, mask) * ((p.tlist[1] - p.tlist[0]) * 0.5);
    }
    const int rr = own ? r : N - 1;
    for (int n = 0; n < NT; ++n) {
        sw_issue<W>(ring + ((n + WARP_D - 1) % WARP_D) * stage, Ug + (size_t)(n + WARP_D - 1) * ustride, N, NS, r,
                    n + WARP_D - 1 < NT);
        cp_async_wait<WARP_D - 1>();
        __syncwarp(mask);
        const cplx* u = ring + (n % WARP_D) * stage + (size_t)rr * NS;
        cplx a0 = mk(0.0, 0.0), a1 = mk(0.0, 0.0);
        int j = 0;
        for (; j + 1 < N; j += 2) {
            const cplx x0 = sub_bcast<W>(psi, j, mask), x1 = sub_bcast<W>(psi, j + 1, mask);
            cfma(a0, u[j], x0);
            cfma(a1, u[j + 1], x1);
        }
        if (j < N) cfma(a0, u[j], sub_bcast<W>(psi, j, mask));
        psi = own ? cadd(a0, a1) : mk(0.0, 0.0);
        __syncwarp(mask);   // ring slot may be overwritten by the next issue
        if (own) st_cs(&p.psi[((size_t)(n + 1) * K + k) * N + r], psi);
        if (gb) {
            const int ntl = n + 1;
            const double w = (ntl < NT) ? 0.5 * (p.tlist[ntl + 1] - p.tlist[ntl - 1])
                                        : 0.5 * (p.tlist[NT] - p.tlist[NT - 1]);
            const cplx t = sw_rowdot_global<W>(Dk, N, r, psi, mask);
            jb = fma(sub_sum<W>(own ? psi.x * t.x + psi.y * t.y : 0.0, mask), w, jb);
        }
    }
    cplx tg = own ? p.tgt[(size_t)k * N + r] : mk(0.0, 0.0);
    cplx acc = mk(0.0, 0.0);
    cfmac(acc, tg, psi);
    acc.x = sub_sum<W>(acc.x, mask);
    acc.y = sub_sum<W>(acc.y, mask);
    if (r == 0) { p.tau[k] = acc; p.jb[k] = jb; }
}

template <int W>
__global__ void warp_backward(DevP p, int NS, const cplx* __restrict__ chi_host) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N = p.N, NN = N * N, G = p.G, NT = p.NT, K = p.K;
    const int spb = blockDim.x / W;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const int k = blockIdx.x * spb + sub;
    if (k >= K) return;
    const int g = p.gen[k];
    const size_t stage = (size_t)N * NS;
    cplx* ring = reinterpret_cast<cplx*>(smraw) + (size_t)sub * WARP_D * stage;
    const cplx* Ug = p.U + (size_t)g * NN;
    const size_t ustride = (size_t)G * NN;
#pragma unroll
    for (int q = 0; q < WARP_D - 1; ++q)
        sw_issue<W>(ring + (q % WARP_D) * stage, Ug + (size_t)max(NT - 1 - q, 0) * ustride, N, NS, r, q < NT);
    const bool own = r < N;
    const int rr = own ? r : N - 1;
    const bool gb = p.gb_kind != 0 && p.lambda_b != 0.0;
    const cplx* Dk = gb ? p.D + (p.gb_nD == 1 ? 0 : (size_t)k * NN) : nullptr;

    cplx x;
    if (chi_host) x = own ? chi_host[(size_t)k * N + r] : mk(0.0, 0.0);
    else {
        const double w = p.w ? p.w[k] : 1.0;
        const double Kg = (double)p.Kglobal;
        cplx c;
        if (p.functional == 0) c = mk(w * p.sums[0] / (Kg * Kg), w * p.sums[1] / (Kg * Kg));
        else if (p.functional == 1) c = mk(w / (2.0 * Kg), 0.0);
        else { cplx t = p.tau[k]; c = mk(w * t.x / Kg, w * t.y / Kg); }
        x = own ? cmul(c, p.tgt[(size_t)k * N + r]) : mk(0.0, 0.0);
    }
    if (gb) {
        const cplx pT = own ? p.psi[((size_t)NT * K + k) * N + r] : mk(0.0, 0.0);
        const cplx t = sw_rowdot_global<W>(Dk, N, r, pT, mask);
        const double f = p.lambda_b * (p.tlist[NT] - p.tlist[NT - 1]) * 0.5;
        if (own) { x.x = fma(-f, t.x, x.x); x.y = fma(-f, t.y, x.y); }
    }
    double rho = sqrt(sub_sum<W>(own ? cnorm2(x) : 0.0, mask));
    if (!(rho >= p.chi_min_norm)) {
        if (r == 0 && atomicCAS(&p.flags->chi_bad_k, 0, k + 1) == 0) p.flags->chi_bad_rho = rho;
        rho = 1.0;
    }
    x = cscale(x, 1.0 / rho);
    if (r == 0) p.rho[k] = rho;
    if (own) p.chiT[(size_t)k * N + r] = x;

    for (int q = 0; q < NT; ++q) {
        const int n = NT - 1 - q;
        sw_issue<W>(ring + ((q + WARP_D - 1) % WARP_D) * stage, Ug + (size_t)max(NT - 1 - (q + WARP_D - 1), 0) * ustride,
                    N, NS, r, q + WARP_D - 1 < NT);
        cplx pp = mk(0.0, 0.0);
        if (gb && n > 0 && own) pp = ld_cs(&p.psi[((size_t)n * K + k) * N + r]);
        if (own) st_cs(&p.chi[((size_t)(n + 1) * K + k) * N + r], x);
        cp_async_wait<WARP_D - 1>();
        __syncwarp(mask);
        const cplx* u = ring + (q % WARP_D) * stage + rr;    // column rr: u[j*NS]
        cplx a0 = mk(0.0, 0.0), a1 = mk(0.0, 0.0);
        int j = 0;
        for (; j + 1 < N; j += 2) {
            const cplx x0 = sub_bcast<W>(x, j, mask), x1 = sub_bcast<W>(x, j + 1, mask);
            cfmac(a0, u[(size_t)j * NS], x0);
            cfmac(a1, u[(size_t)(j + 1) * NS], x1);
        }
        if (j < N) cfmac(a0, u[(size_t)j * NS], sub_bcast<W>(x, j, mask));
        x = own ? cadd(a0, a1) : mk(0.0, 0.0);
        __syncwarp(mask);
        if (gb && n > 0) {
            const cplx t = sw_rowdot_global<W>(Dk, N, r, pp, mask);
            const double f = p.lambda_b * 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]) / rho;
            if (own) { x.x = fma(-f, t.x, x.x); x.y = fma(-f, t.y, x.y); }
        }
    }
}

// ---------------------------------------------------------------------------
// Phase C: sub-warp per (k,n); H_n in shared memory, lane i owns row i of the
// block vectors [chi'_1 .. chi'_LC, chi].
// ---------------------------------------------------------------------------
template <int W, int LC>
__global__ void warp_gradient(DevP p, int l0, int spb) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N = p.N, NN = N * N, NT = p.NT, K = p.K;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const long long unit = (long long)blockIdx.x * spb + sub;
    if (unit >= (long long)K * NT) return;
    const int k = (int)(unit % K), n = (int)(unit / K);
    const int g = p.gen[k];
    cplx* Hs = reinterpret_cast<cplx*>(smraw) + (size_t)sub * NN;
    const double dt = p.tlist[n + 1] - p.tlist[n];
    const cplx* H0 = p.H0 + (size_t)g * NN;
    const cplx* Hc = p.Hc + (size_t)g * p.L * NN;
    for (int e = r; e < NN; e += W) {
        cplx h = __ldg(&H0[e]);
        for (int l = 0; l < p.L; ++l) {
            double a = p.eps[l * NT + n];
            if (p.shape) a *= p.shape[l * NT + n];
            cfmar(h, a, __ldg(&Hc[(size_t)l * NN + e]));
        }
        Hs[e] = h;
    }
    __syncwarp(mask);
    const bool own = r < N;
    const int rr = own ? r : N - 1;
    double cs = 0.0;
    if (own)
        for (int j = 0; j < N; ++j) cs += cabs1(Hs[j * N + r]);
    int m, s;
    {
        const double nrm = dt * sub_max<W>(cs, mask);
        if (p.grad_method == 0) vec_plan(nrm, m, s);
        else { s = 0; m = p.taylor_max_order; }
    }
    const double sc = dt * ldexp(1.0, -s);
    double scl[LC];
#pragma unroll
    for (int l = 0; l < LC; ++l) scl[l] = p.dshape ? sc * p.dshape[(l0 + l) * NT + n] : sc;
    const cplx* Hcl = Hc + (size_t)l0 * NN;

    cplx asum = own ? ld_cs(&p.chi[((size_t)(n + 1) * K + k) * N + r]) : mk(0.0, 0.0);
    cplx bsum[LC];
#pragma unroll
    for (int l = 0; l < LC; ++l) bsum[l] = mk(0.0, 0.0);
    const bool taylor = p.grad_method != 0;
    bool converged = !taylor || !p.taylor_check;
    double rlast = 0.0;
    bool done[LC];
#pragma unroll
    for (int l = 0; l < LC; ++l) done[l] = false;
    for (int subst = 0; subst < (1 << s); ++subst) {
        cplx ta = asum, tb[LC];
#pragma unroll
        for (int l = 0; l < LC; ++l) tb[l] = bsum[l];
        for (int j = 1; j <= m; ++j) {
            const double inv = 1.0 / (double)j;
            cplx na = mk(0.0, 0.0), nb[LC];
#pragma unroll
            for (int l = 0; l < LC; ++l) nb[l] = mk(0.0, 0.0);
            for (int c = 0; c < N; ++c) {
                // Abar_{r,c} = i*sc*conj(H_{c,r}) ; Ebar_l similarly from Hc_l
                const cplx h = Hs[c * N + rr];
                const cplx a = mk(sc * h.y, sc * h.x);
                const cplx tac = sub_bcast<W>(ta, c, mask);
                cfma(na, a, tac);
#pragma unroll
                for (int l = 0; l < LC; ++l) {
                    const cplx tbc = sub_bcast<W>(tb[l], c, mask);
                    cfma(nb[l], a, tbc);
                    const cplx e = __ldg(&Hcl[(size_t)l * NN + c * N + rr]);
                    cfma(nb[l], mk(scl[l] * e.y, scl[l] * e.x), tac);
                }
            }
            ta = own ? cscale(na, inv) : mk(0.0, 0.0);
            asum = cadd(asum, ta);
            bool all_done = true;
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                tb[l] = own ? cscale(nb[l], inv) : mk(0.0, 0.0);
                if (!done[l]) bsum[l] = cadd(bsum[l], tb[l]);
                if (taylor && p.taylor_check && j >= 2 && !done[l]) {   // per-control early return, optimize.jl:633-638
                    rlast = sqrt(sub_sum<W>(cnorm2(tb[l]), mask));
                    if (rlast < p.taylor_tol) done[l] = true;
                }
                all_done = all_done && done[l];
            }
            if (taylor && p.taylor_check && j >= 2 && all_done) { converged = true; break; }
        }
    }
    if (!converged && p.taylor_max_order > 1 && r == 0) {
        if (atomicExch(&p.flags->taylor_fail, 1) == 0) p.flags->taylor_r = rlast;
    }
    const double rho = p.rho[k];
    const cplx ps = own ? ld_cs(&p.psi[((size_t)n * K + k) * N + r]) : mk(0.0, 0.0);
#pragma unroll
    for (int l = 0; l < LC; ++l) {
        cplx acc = mk(0.0, 0.0);
        cfmac(acc, bsum[l], ps);
        acc.x = rho * sub_sum<W>(acc.x, mask);
        acc.y = rho * sub_sum<W>(acc.y, mask);
        if (r == 0) {
            p.partial[(size_t)k * p.L * NT + (size_t)(l0 + l) * NT + n] = acc.x;
            if (p.taugrads) p.taugrads[((size_t)k * p.L + (l0 + l)) * NT + n] = acc;
        }
    }
}

__global__ void warp_gather_states_k(const cplx* __restrict__ psi, cplx* __restrict__ out, int K, int N, int NT, int k) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (NT + 1) * N) out[idx] = psi[((size_t)(idx / N) * K + k) * N + (idx % N)];
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#define WARP_SWITCH(Wv, ...)                                        \
    switch (Wv) {                                                   \
        case 8: { constexpr int WW = 8; __VA_ARGS__; } break;       \
        case 16: { constexpr int WW = 16; __VA_ARGS__; } break;     \
        default: { constexpr int WW = 32; __VA_ARGS__; } break;     \
    }

inline int warp_setup(WarpPlan& wp, DevP& p, const grape_b200_problem* d, std::vector<void*>& allocs, std::string& err) {
    const int K = p.K, N = p.N, L = p.L, NT = p.NT, G = p.G, NN = N * N;
    wp.W = N <= 8 ? 8 : (N <= 16 ? 16 : 32);
    wp.NS = N | 1;
    auto up = [&](const std::vector<cplx>& b, const cplx** dst) -> int {
        void* q = nullptr;
        if (cudaMalloc(&q, b.size() * sizeof(cplx)) != cudaSuccess) { err = "cudaMalloc failed (warp path)"; return GRAPE_B200_ECUDA; }
        allocs.push_back(q);
        if (cudaMemcpy(q, b.data(), b.size() * sizeof(cplx), cudaMemcpyHostToDevice) != cudaSuccess) { err = "cudaMemcpy failed"; return GRAPE_B200_ECUDA; }
        *dst = static_cast<const cplx*>(q);
        return 0;
    };
    auto al = [&](cplx** dst, size_t n) -> int {
        void* q = nullptr;
        if (cudaMalloc(&q, n * sizeof(cplx)) != cudaSuccess) { err = "cudaMalloc failed (warp path)"; return GRAPE_B200_ECUDA; }
        allocs.push_back(q);
        *dst = static_cast<cplx*>(q);
        return 0;
    };
    // column-major ABI -> row-major device
    auto to_rowmajor = [&](const double* src, size_t nmat) {
        std::vector<cplx> b(nmat * NN);
        for (size_t mtx = 0; mtx < nmat; ++mtx)
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) {
                    const double* s = src + 2 * (mtx * NN + (size_t)j * N + i);
                    b[mtx * NN + (size_t)i * N + j] = mk(s[0], s[1]);
                }
        return b;
    };
    int rc;
    if ((rc = up(to_rowmajor(d->H0, G), &p.H0))) return rc;
    if ((rc = up(to_rowmajor(d->Hc, (size_t)G * L), &p.Hc))) return rc;
    if (p.gb_kind && (rc = up(to_rowmajor(d->gb_D, p.gb_nD), &p.D))) return rc;
    {
        std::vector<cplx> b((size_t)K * N);
        for (size_t e = 0; e < b.size(); ++e) b[e] = mk(d->psi0[2 * e], d->psi0[2 * e + 1]);
        if ((rc = up(b, &p.psi0))) return rc;
        for (size_t e = 0; e < b.size(); ++e) b[e] = mk(d->tgt[2 * e], d->tgt[2 * e + 1]);
        if ((rc = up(b, &p.tgt))) return rc;
    }
    if ((rc = al(&p.U, (size_t)NT * G * NN))) return rc;
    if ((rc = al(&p.psi, (size_t)(NT + 1) * K * N))) return rc;
    if ((rc = al(&p.chi, (size_t)(NT + 1) * K * N))) return rc;
    p.KB = K;
    {
        void* q = nullptr;
        if (cudaMalloc(&q, (size_t)K * L * NT * sizeof(double)) != cudaSuccess) { err = "cudaMalloc failed (warp path)"; return GRAPE_B200_ECUDA; }
        allocs.push_back(q);
        p.partial = static_cast<double*>(q);
    }
    const size_t perA = (size_t)4 * NN * sizeof(cplx);
    wp.spbA = (int)std::max<size_t>(1, std::min<size_t>(128 / wp.W, (size_t)(200 * 1024) / perA));
    wp.smemA = perA * wp.spbA;
    wp.smemB = (size_t)WARP_D * N * wp.NS * sizeof(cplx) * (32 / wp.W);
    wp.spbC = 128 / wp.W;
    wp.smemC = (size_t)NN * sizeof(cplx) * wp.spbC;
    cudaError_t e = cudaSuccess;
#define SETATTR(fn, bytes) if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
    WARP_SWITCH(wp.W,
        SETATTR(warp_form_U<WW>, wp.smemA); SETATTR(warp_forward<WW>, wp.smemB); SETATTR(warp_backward<WW>, wp.smemB);
        SETATTR((warp_gradient<WW, 1>), wp.smemC); SETATTR((warp_gradient<WW, 2>), wp.smemC); SETATTR((warp_gradient<WW, 4>), wp.smemC))
#undef SETATTR
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(e); return GRAPE_B200_ECUDA; }
    return 0;
}

inline void warp_run_formU(WarpPlan& wp, const DevP& p, cudaStream_t st, int64_t& launches) {
    const long long units = (long long)p.G * p.NT;
    const unsigned blocks = (unsigned)((units + wp.spbA - 1) / wp.spbA);
    const int threads = wp.spbA * wp.W;
    WARP_SWITCH(wp.W, warp_form_U<WW><<<blocks, threads, wp.smemA, st>>>(p, wp.spbA))
    launches++;
}
inline void warp_run_forward(WarpPlan& wp, const DevP& p, cudaStream_t st, int64_t& launches) {
    const int spb = 32 / wp.W;
    const unsigned blocks = (unsigned)((p.K + spb - 1) / spb);
    WARP_SWITCH(wp.W, warp_forward<WW><<<blocks, 32, wp.smemB, st>>>(p, wp.NS))
    launches++;
}
inline void warp_run_backward(WarpPlan& wp, const DevP& p, const cplx* chi_host, cudaStream_t st, int64_t& launches) {
    const int spb = 32 / wp.W;
    const unsigned blocks = (unsigned)((p.K + spb - 1) / spb);
    WARP_SWITCH(wp.W, warp_backward<WW><<<blocks, 32, wp.smemB, st>>>(p, wp.NS, chi_host))
    launches++;
}
template <int W>
inline void warp_gradient_launch(WarpPlan& wp, const DevP& p, cudaStream_t st, int64_t& launches) {
    const long long units = (long long)p.K * p.NT;
    const unsigned blocks = (unsigned)((units + wp.spbC - 1) / wp.spbC);
    int l0 = 0;
    while (l0 < p.L) {
        const int rem = p.L - l0;
        if (rem >= 4) { warp_gradient<W, 4><<<blocks, 128, wp.smemC, st>>>(p, l0, wp.spbC); l0 += 4; }
        else if (rem >= 2) { warp_gradient<W, 2><<<blocks, 128, wp.smemC, st>>>(p, l0, wp.spbC); l0 += 2; }
        else { warp_gradient<W, 1><<<blocks, 128, wp.smemC, st>>>(p, l0, wp.spbC); l0 += 1; }
        launches++;
    }
}
inline void warp_run_gradient(WarpPlan& wp, const DevP& p, cudaStream_t st, int64_t& launches) {
    WARP_SWITCH(wp.W, warp_gradient_launch<WW>(wp, p, st, launches))
}
inline void warp_gather_final(WarpPlan&, const DevP& p, cplx* out, cudaStream_t st, int64_t&) {
    cudaMemcpyAsync(out, p.psi + (size_t)p.NT * p.K * p.N, sizeof(cplx) * (size_t)p.K * p.N, cudaMemcpyDeviceToDevice, st);
}
inline void warp_gather_states(WarpPlan&, const DevP& p, int k, cplx* out, cudaStream_t st, int64_t& launches) {
    const int cnt = (p.NT + 1) * p.N;
    warp_gather_states_k<<<(cnt + 255) / 256, 256, 0, st>>>(p.psi, out, p.K, p.N, p.NT, k);
    launches++;
}
