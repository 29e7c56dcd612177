// warp_seg.cuh -- time-segmented ("scan") schedule of the 5 <= N <= 32 path.
//
// Same decomposition as small_seg.cuh: the NT-step chains of warp_n.cuh are cut into NSEG segments
// of S steps.  Segment propagators P_{g,seg} are products of the U_{g,n}; a chain of only NSEG
// steps gives the states at the segment boundaries; the interior of every segment is then filled in
// parallel over (trajectory, segment).  Both Psi (fw_storage, reference src/workspace.jl:215) and
// chi are stored for every time point, and the gradient contraction stays the fully parallel
// warp_gradient kernel of warp_n.cuh.
//
// One sub-warp (W = 8, 16 or 32 lanes) per unit; lane r owns element r of the state and keeps row r
// (forward) or column r (backward, U^dagger) of the step's matrix in registers, loaded one step ahead.
#pragma once
#include "warp_n.cuh"

struct WarpSegArgs {
    int S, NSEG;
    cplx* Pseg;   // [NSEG][G][N*N] row-major
};

// ---------------------------------------------------------------------------
// segment propagators: P = U_{n1-1} ... U_{n0}; sub-warp per (g, seg); matrices in shared memory
// ---------------------------------------------------------------------------
template <int W>
__global__ void warp_segprod(DevP p, WarpSegArgs a, int spb) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N = p.N, NN = N * N, G = p.G, NT = p.NT;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const long long unit = (long long)blockIdx.x * spb + sub;
    if (unit >= (long long)G * a.NSEG) return;
    const int g = (int)(unit % G), seg = (int)(unit / G);
    const int n0 = seg * a.S, n1 = min(NT, n0 + a.S);
    cplx* P = reinterpret_cast<cplx*>(smraw) + (size_t)sub * 3 * NN;
    cplx* T = P + NN;
    cplx* Us = T + NN;
    const cplx* Ug = p.U + (size_t)g * NN;
    const size_t ustride = (size_t)G * NN;
    for (int e = r; e < NN; e += W) P[e] = Ug[(size_t)n0 * ustride + e];
    __syncwarp(mask);
    for (int n = n0 + 1; n < n1; ++n) {
        for (int e = r; e < NN; e += W) Us[e] = Ug[(size_t)n * ustride + e];
        __syncwarp(mask);
        sw_matmul<W>(T, Us, P, N, r, mask);   // T = U_n * P
        cplx* tmp = P; P = T; T = tmp;
    }
    cplx* o = a.Pseg + ((size_t)seg * G + g) * NN;
    for (int e = r; e < NN; e += W) o[e] = P[e];
}

// y_r = sum_j M[r][j] x_j (row r in u[]) or sum_j conj(M[j][r]) x_j (column r in u[])
template <int W, bool ADJ>
GB_D cplx wseg_apply(const cplx (&u)[W], cplx x, int N, unsigned mask) {
    cplx a0 = mk(0.0, 0.0), a1 = mk(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < W; j += 2) {
        if (j < N) {
            const cplx x0 = sub_bcast<W>(x, j, mask);
            if (ADJ) cfmac(a0, u[j], x0); else cfma(a0, u[j], x0);
        }
        if (j + 1 < N) {
            const cplx x1 = sub_bcast<W>(x, j + 1, mask);
            if (ADJ) cfmac(a1, u[j + 1], x1); else cfma(a1, u[j + 1], x1);
        }
    }
    return cadd(a0, a1);
}

// matrix for the current step: taken from the one-step-ahead buffer (W <= 16) or loaded now (W = 32,
// where two W-element register rows would not fit)
template <int W, bool ADJ, int WA>
GB_D void wseg_load(cplx (&u)[WA], const cplx* __restrict__ M, int N, int rr) {
    if (WA == W) {
#pragma unroll
        for (int j = 0; j < WA; ++j)
            if (j < N) u[j] = ADJ ? __ldg(&M[(size_t)j * N + rr]) : __ldg(&M[(size_t)rr * N + j]);
    }
}

// ---------------------------------------------------------------------------
// forward: CHAIN over the segment propagators (unit = k) or FILL inside the segments (unit = (k, seg))
// ---------------------------------------------------------------------------
template <int W, bool FILL>
__global__ void __launch_bounds__(128) warp_seg_fwd(DevP p, WarpSegArgs a) {
    const int N = p.N, NN = N * N, G = p.G, NT = p.NT, K = p.K;
    const int spb = blockDim.x / W;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const long long unit = (long long)blockIdx.x * spb + sub;
    const long long units = FILL ? (long long)K * a.NSEG : (long long)K;
    if (unit >= units) return;
    const int k = (int)(unit % K), seg = (int)(unit / K);
    const int g = p.gen[k];
    const bool own = r < N;
    const int rr = own ? r : N - 1;
    const size_t mstride = (size_t)G * NN;
    constexpr bool PF = W <= 16;
    cplx u[W], un[PF ? W : 1];
    if (FILL) {
        const int n0 = seg * a.S, n1 = min(NT, n0 + a.S);
        if (n1 - n0 < 2) return;
        const cplx* Ug = p.U + (size_t)g * NN;
        cplx x = own ? p.psi[((size_t)n0 * K + k) * N + r] : mk(0.0, 0.0);
        if (PF) wseg_load<W, false>(un, Ug + (size_t)n0 * mstride, N, rr);
        for (int n = n0; n < n1 - 1; ++n) {
            if (PF) {
#pragma unroll
                for (int j = 0; j < W; ++j) u[j] = un[PF ? j : 0];
                if (n + 1 < n1 - 1) wseg_load<W, false>(un, Ug + (size_t)(n + 1) * mstride, N, rr);
            } else wseg_load<W, false>(u, Ug + (size_t)n * mstride, N, rr);
            x = wseg_apply<W, false>(u, x, N, mask);
            if (!own) x = mk(0.0, 0.0);
            if (own) st_cs(&p.psi[((size_t)(n + 1) * K + k) * N + r], x);
        }
    } else {
        const cplx* Pg = a.Pseg + (size_t)g * NN;
        cplx x = own ? p.psi0[(size_t)k * N + r] : mk(0.0, 0.0);
        if (own) p.psi[(size_t)k * N + r] = x;
        if (PF) wseg_load<W, false>(un, Pg, N, rr);
        for (int q = 0; q < a.NSEG; ++q) {
            if (PF) {
#pragma unroll
                for (int j = 0; j < W; ++j) u[j] = un[PF ? j : 0];
                if (q + 1 < a.NSEG) wseg_load<W, false>(un, Pg + (size_t)(q + 1) * mstride, N, rr);
            } else wseg_load<W, false>(u, Pg + (size_t)q * mstride, N, rr);
            x = wseg_apply<W, false>(u, x, N, mask);
            if (!own) x = mk(0.0, 0.0);
            const int nb = min(NT, (q + 1) * a.S);
            if (own) st_cs(&p.psi[((size_t)nb * K + k) * N + r], x);
        }
        const cplx tg = own ? p.tgt[(size_t)k * N + r] : mk(0.0, 0.0);
        cplx acc = mk(0.0, 0.0);
        cfmac(acc, tg, x);
        acc.x = sub_sum<W>(acc.x, mask);
        acc.y = sub_sum<W>(acc.y, mask);
        if (r == 0) { p.tau[k] = acc; p.jb[k] = 0.0; }
    }
}

// ---------------------------------------------------------------------------
// backward: chi index convention of warp_n.cuh: chi[t] = chi_k at time point t (t = 1..NT)
// ---------------------------------------------------------------------------
template <int W, bool FILL>
__global__ void __launch_bounds__(128) warp_seg_bwd(DevP p, WarpSegArgs a, const cplx* __restrict__ chi_host) {
    const int N = p.N, NN = N * N, G = p.G, NT = p.NT, K = p.K;
    const int spb = blockDim.x / W;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const long long unit = (long long)blockIdx.x * spb + sub;
    const long long units = FILL ? (long long)K * a.NSEG : (long long)K;
    if (unit >= units) return;
    const int k = (int)(unit % K), seg = (int)(unit / K);
    const int g = p.gen[k];
    const bool own = r < N;
    const int rr = own ? r : N - 1;
    const size_t mstride = (size_t)G * NN;
    constexpr bool PF = W <= 16;
    cplx u[W], un[PF ? W : 1];
    if (FILL) {
        const int n0 = seg * a.S, n1 = min(NT, n0 + a.S);
        if (n1 - n0 < 2) return;
        const cplx* Ug = p.U + (size_t)g * NN;
        cplx x = own ? p.chi[((size_t)n1 * K + k) * N + r] : mk(0.0, 0.0);
        if (PF) wseg_load<W, true>(un, Ug + (size_t)(n1 - 1) * mstride, N, rr);
        for (int n = n1 - 1; n > n0; --n) {
            if (PF) {
#pragma unroll
                for (int j = 0; j < W; ++j) u[j] = un[PF ? j : 0];
                if (n - 1 > n0) wseg_load<W, true>(un, Ug + (size_t)(n - 1) * mstride, N, rr);
            } else wseg_load<W, true>(u, Ug + (size_t)n * mstride, N, rr);
            x = wseg_apply<W, true>(u, x, N, mask);
            if (!own) x = mk(0.0, 0.0);
            if (own) st_cs(&p.chi[((size_t)n * K + k) * N + r], x);
        }
    } else {
        const cplx* Pg = a.Pseg + (size_t)g * NN;
        if (PF && a.NSEG > 1) wseg_load<W, true>(un, Pg + (size_t)(a.NSEG - 1) * mstride, N, rr);
        // boundary condition chi_k(T) (reference src/optimize.jl:845-869)
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
        double rho = sqrt(sub_sum<W>(own ? cnorm2(x) : 0.0, mask));
        if (!(rho >= p.chi_min_norm)) {
            if (r == 0 && atomicCAS(&p.flags->chi_bad_k, 0, k + 1) == 0) p.flags->chi_bad_rho = rho;
            rho = 1.0;
        }
        x = cscale(x, 1.0 / rho);
        if (r == 0) p.rho[k] = rho;
        if (own) {
            p.chiT[(size_t)k * N + r] = x;
            p.chi[((size_t)NT * K + k) * N + r] = x;
        }
        for (int q = a.NSEG - 1; q >= 1; --q) {
            if (PF) {
#pragma unroll
                for (int j = 0; j < W; ++j) u[j] = un[PF ? j : 0];
                if (q - 1 >= 1) wseg_load<W, true>(un, Pg + (size_t)(q - 1) * mstride, N, rr);
            } else wseg_load<W, true>(u, Pg + (size_t)q * mstride, N, rr);
            x = wseg_apply<W, true>(u, x, N, mask);
            if (!own) x = mk(0.0, 0.0);
            if (own) st_cs(&p.chi[((size_t)(q * a.S) * K + k) * N + r], x);
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
inline int warp_seg_setup(WarpSegArgs& a, const WarpPlan& wp, const DevP& p, std::vector<void*>& allocs, std::string& err) {
    int S = (int)std::ceil(std::sqrt((double)p.NT));
    if (const char* e = getenv("GRAPE_B200_SEG_S")) S = atoi(e);
    a.S = S < 2 ? 2 : (S > 128 ? 128 : S);
    a.NSEG = (p.NT + a.S - 1) / a.S;
    void* q = nullptr;
    if (cudaMalloc(&q, (size_t)a.NSEG * p.G * p.N * p.N * sizeof(cplx)) != cudaSuccess) { err = "cudaMalloc failed (warp seg)"; return GRAPE_B200_ECUDA; }
    allocs.push_back(q);
    a.Pseg = static_cast<cplx*>(q);
    cudaError_t e = cudaSuccess;
    const size_t smem = (size_t)3 * p.N * p.N * sizeof(cplx) * (128 / wp.W);
    WARP_SWITCH(wp.W, e = cudaFuncSetAttribute(warp_segprod<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute failed (warp seg): ") + cudaGetErrorString(e); return GRAPE_B200_ECUDA; }
    return 0;
}

inline void warp_seg_run_prod(const WarpSegArgs& a, const WarpPlan& wp, const DevP& p, cudaStream_t st, int64_t& launches) {
    const int spb = 128 / wp.W;
    const long long units = (long long)p.G * a.NSEG;
    const size_t smem = (size_t)3 * p.N * p.N * sizeof(cplx) * spb;
    WARP_SWITCH(wp.W, warp_segprod<WW><<<(unsigned)((units + spb - 1) / spb), 128, smem, st>>>(p, a, spb))
    launches++;
}
inline void warp_seg_run_forward(const WarpSegArgs& a, const WarpPlan& wp, const DevP& p, bool fill, cudaStream_t st, int64_t& launches) {
    const int spb = 128 / wp.W;
    {
        const long long units = p.K;
        WARP_SWITCH(wp.W, (warp_seg_fwd<WW, false><<<(unsigned)((units + spb - 1) / spb), 128, 0, st>>>(p, a)))
        launches++;
    }
    if (fill) {
        const long long units = (long long)p.K * a.NSEG;
        WARP_SWITCH(wp.W, (warp_seg_fwd<WW, true><<<(unsigned)((units + spb - 1) / spb), 128, 0, st>>>(p, a)))
        launches++;
    }
}
inline void warp_seg_run_fill(const WarpSegArgs& a, const WarpPlan& wp, const DevP& p, cudaStream_t st, int64_t& launches) {
    const int spb = 128 / wp.W;
    const long long units = (long long)p.K * a.NSEG;
    WARP_SWITCH(wp.W, (warp_seg_fwd<WW, true><<<(unsigned)((units + spb - 1) / spb), 128, 0, st>>>(p, a)))
    launches++;
}
inline void warp_seg_run_backward(const WarpSegArgs& a, const WarpPlan& wp, const DevP& p, const cplx* chi_host, cudaStream_t st, int64_t& launches) {
    const int spb = 128 / wp.W;
    {
        const long long units = p.K;
        WARP_SWITCH(wp.W, (warp_seg_bwd<WW, false><<<(unsigned)((units + spb - 1) / spb), 128, 0, st>>>(p, a, chi_host)))
    }
    {
        const long long units = (long long)p.K * a.NSEG;
        WARP_SWITCH(wp.W, (warp_seg_bwd<WW, true><<<(unsigned)((units + spb - 1) / spb), 128, 0, st>>>(p, a, chi_host)))
    }
    launches += 2;
}

