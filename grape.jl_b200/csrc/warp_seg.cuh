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

//
// Round 2, Hermitian generators (every U is unitary): SCAN schedule.  The two boundary chains (NSEG dependent
// mat-vecs each, one sub-warp per trajectory: 2 x 36 us of the 283 us C2 gradient) are replaced by the prefix
// products Q_seg = P_seg .. P_0 of the segment propagators -- a Kogge-Stone scan over the segments, ceil(log2 NSEG)
// levels of independent N x N products, one short kernel per level -- after which every boundary state is one or two
// mat-vecs, independent of all the others:
//     Psi_k(end of seg) = Q_seg Psi_k(0),   chi_k(end of seg) = Q_seg Q_last^dagger chi_k(T).
// Without the chains short segments only cost scan levels: S = NT / 256 instead of sqrt(NT), so the segment products
// and both fills are 8 instead of 45 dependent steps deep on C2.
struct WarpSegArgs {
    int S, NSEG;
    cplx* Pseg;   // [NSEG][G][N*N] row-major
    int scan;     // scan schedule (Hermitian generators)
    int pf;       // warp_segprod: propagators double-buffered through cp.async
    int radix;    // factors per scan level (2 or 4)
    cplx* Qa;     // ping-pong buffers of the scan, same layout as Pseg
    cplx* Qb;
    const cplx* Q;   // where the prefix products end up (Pseg, Qa or Qb: fixed by the number of levels)
};

// ---------------------------------------------------------------------------
// segment propagators: P = U_{n1-1} ... U_{n0}; sub-warp per (g, seg); matrices in shared memory
// ---------------------------------------------------------------------------
template <int W>
__global__ void warp_segprod(DevP p, WarpSegArgs a, int spb, int pf) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N = p.N, NN = N * N, G = p.G, NT = p.NT;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const long long unit = (long long)blockIdx.x * spb + sub;
    if (unit >= (long long)G * a.NSEG) return;
    const int g = (int)(unit % G), seg = (int)(unit / G);
    const int n0 = seg * a.S, n1 = min(NT, n0 + a.S);
    cplx* P = reinterpret_cast<cplx*>(smraw) + (size_t)sub * (3 + pf) * NN;
    cplx* T = P + NN;
    cplx* Us = T + NN;
    cplx* Us2 = Us + NN;   // pf: second buffer, the next step's propagator arrives (cp.async) while this one is multiplied
    const cplx* Ug = p.U + (size_t)g * NN;
    const size_t ustride = (size_t)G * NN;
    for (int e = r; e < NN; e += W) P[e] = Ug[(size_t)n0 * ustride + e];
    if (pf) {
        if (n0 + 1 < n1)
            for (int e = r; e < NN; e += W) cp_async16(Us + e, Ug + (size_t)(n0 + 1) * ustride + e);
        cp_async_commit();
    }
    __syncwarp(mask);
    for (int n = n0 + 1; n < n1; ++n) {
        cplx* cur = Us;
        if (pf) {
            cur = ((n - n0) & 1) ? Us : Us2;
            cplx* nxt = ((n - n0) & 1) ? Us2 : Us;
            cp_async_wait<0>();
            __syncwarp(mask);
            if (n + 1 < n1)
                for (int e = r; e < NN; e += W) cp_async16(nxt + e, Ug + (size_t)(n + 1) * ustride + e);
            cp_async_commit();
        } else {
            for (int e = r; e < NN; e += W) Us[e] = Ug[(size_t)n * ustride + e];
            __syncwarp(mask);
        }
        sw_matmul<W>(T, cur, P, N, r, mask);   // T = U_n * P
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
// scan schedule
// ---------------------------------------------------------------------------
// one level of the inclusive prefix products over the segments (Kogge-Stone with radix R = 2 or 4):
//     x_seg <- x_seg x_{seg-off} [x_{seg-2 off} x_{seg-3 off}],   off = 1, R, R^2, ..
// from `in` to `out`; sub-warp per (g, seg) over the whole grid, operands staged in shared memory.  One launch per level
// inside the call's CUDA graph.  Measured on C2 (profiles/r2_s23_*, r2_s24_*): one block per generator with operands in
// global memory took 7 us per level (L2 round trips inside the product loop), a radix-2 level launched grid-wide 5 us --
// mostly launch latency, hence radix 4: half the levels, three short products each.
template <int W>
__global__ void warp_scan_level(DevP p, WarpSegArgs a, const cplx* __restrict__ in, cplx* __restrict__ out, int off, int R, int spb) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int N = p.N, NN = N * N, G = p.G;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const long long unit = (long long)blockIdx.x * spb + sub;
    if (unit >= (long long)G * a.NSEG) return;
    const int g = (int)(unit % G), seg = (int)(unit / G);
    const cplx* x = in + ((size_t)seg * G + g) * NN;
    cplx* o = out + ((size_t)seg * G + g) * NN;
    if (seg < off) {
        for (int e = r; e < NN; e += W) o[e] = x[e];
        return;
    }
    cplx* buf = reinterpret_cast<cplx*>(smraw) + (size_t)sub * (R + 1) * NN;   // R operands + one spare
    const int cnt = min(R, seg / off + 1);                                      // factors x_seg, x_{seg-off}, ..
    for (int j = 0; j < cnt; ++j) {
        const cplx* y = in + ((size_t)(seg - j * off) * G + g) * NN;
        for (int e = r; e < NN; e += W) buf[(size_t)j * NN + e] = y[e];
    }
    __syncwarp(mask);
    // t = x_seg; t <- t x_{seg - j off}: products go to the spare buffer, then to operand buffers that are done
    const cplx* t = buf;
    for (int j = 1; j < cnt; ++j) {
        cplx* c = j == cnt - 1 ? o : (j == 1 ? buf + (size_t)R * NN : buf + (size_t)(j - 2) * NN);
        sw_matmul<W>(c, t, buf + (size_t)j * NN, N, r, mask);
        t = c;
    }
}

// Psi_k at every segment end from the prefix products, tau_k from the last one; unit = (k, seg)
template <int W>
__global__ void __launch_bounds__(128) warp_scan_bounds_fwd(DevP p, WarpSegArgs a) {
    const int N = p.N, NN = N * N, G = p.G, NT = p.NT, K = p.K;
    const int spb = blockDim.x / W;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const long long unit = (long long)blockIdx.x * spb + sub;
    if (unit >= (long long)K * a.NSEG) return;
    const int k = (int)(unit % K), seg = (int)(unit / K);
    const int g = p.gen[k];
    const bool own = r < N;
    const int rr = own ? r : N - 1;
    cplx u[W];
    wseg_load<W, false>(u, a.Q + ((size_t)seg * G + g) * NN, N, rr);
    const cplx x0 = own ? p.psi0[(size_t)k * N + r] : mk(0.0, 0.0);
    if (seg == 0 && own) p.psi[(size_t)k * N + r] = x0;
    cplx x = wseg_apply<W, false>(u, x0, N, mask);
    if (!own) x = mk(0.0, 0.0);
    const int nb = min(NT, (seg + 1) * a.S);
    if (own) st_cs(&p.psi[((size_t)nb * K + k) * N + r], x);
    if (seg == a.NSEG - 1) {
        const cplx tg = own ? p.tgt[(size_t)k * N + r] : mk(0.0, 0.0);
        cplx acc = mk(0.0, 0.0);
        cfmac(acc, tg, x);
        acc.x = sub_sum<W>(acc.x, mask);
        acc.y = sub_sum<W>(acc.y, mask);
        if (r == 0) { p.tau[k] = acc; p.jb[k] = 0.0; }
    }
}

// chi_k(T) (reference src/optimize.jl:845-869) and chi_k at every segment end: chi(end of seg) = Q_seg Q_last^dagger chi(T);
// unit = (k, seg), the unit of the last segment also writes rho_k, chi_k(T)
template <int W>
__global__ void __launch_bounds__(128) warp_scan_bounds_bwd(DevP p, WarpSegArgs a, const cplx* __restrict__ chi_host) {
    const int N = p.N, NN = N * N, G = p.G, NT = p.NT, K = p.K;
    const int spb = blockDim.x / W;
    const int sub = threadIdx.x / W, r = threadIdx.x % W;
    const unsigned mask = sub_mask<W>();
    const long long unit = (long long)blockIdx.x * spb + sub;
    if (unit >= (long long)K * a.NSEG) return;
    const int k = (int)(unit % K), seg = (int)(unit / K);
    const int g = p.gen[k];
    const bool own = r < N;
    const int rr = own ? r : N - 1;
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
        if (seg == a.NSEG - 1 && r == 0 && atomicCAS(&p.flags->chi_bad_k, 0, k + 1) == 0) p.flags->chi_bad_rho = rho;
        rho = 1.0;
    }
    x = cscale(x, 1.0 / rho);
    if (seg == a.NSEG - 1) {
        if (r == 0) p.rho[k] = rho;
        if (own) {
            p.chiT[(size_t)k * N + r] = x;
            p.chi[((size_t)NT * K + k) * N + r] = x;
        }
        return;   // the whole sub-warp
    }
    cplx u[W];
    wseg_load<W, true>(u, a.Q + ((size_t)(a.NSEG - 1) * G + g) * NN, N, rr);
    cplx y = wseg_apply<W, true>(u, x, N, mask);       // Q_last^dagger chi(T)
    if (!own) y = mk(0.0, 0.0);
    wseg_load<W, false>(u, a.Q + ((size_t)seg * G + g) * NN, N, rr);
    cplx z = wseg_apply<W, false>(u, y, N, mask);
    if (own) st_cs(&p.chi[((size_t)((seg + 1) * a.S) * K + k) * N + r], z);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
inline int warp_seg_setup(WarpSegArgs& a, const WarpPlan& wp, const DevP& p, const grape_b200_problem* d, std::vector<void*>& allocs, std::string& err) {
    int S = (int)std::ceil(std::sqrt((double)p.NT));
    // scan schedule: Hermitian generators only (chi at the segment ends needs Q_seg^{-1} = Q_seg^dagger);
    // GRAPE_B200_WSEG_SCAN=0 keeps the boundary chains
    bool herm = true;
    {
        const int N = p.N;
        auto check = [&](const double* m, size_t count) {
            for (size_t q = 0; q < count && herm; ++q)
                for (int i = 0; i < N && herm; ++i)
                    for (int j = i; j < N; ++j) {
                        const double* x = m + 2 * (q * N * N + (size_t)j * N + i);
                        const double* y = m + 2 * (q * N * N + (size_t)i * N + j);
                        const double tol = 1e-15 * (std::fabs(x[0]) + std::fabs(x[1]) + std::fabs(y[0]) + std::fabs(y[1]));
                        if (std::fabs(x[0] - y[0]) > tol || std::fabs(x[1] + y[1]) > tol) { herm = false; break; }
                    }
        };
        check(d->H0, (size_t)p.G);
        check(d->Hc, (size_t)p.G * p.L);
    }
    a.scan = herm && p.NT >= 16 && !(getenv("GRAPE_B200_WSEG_SCAN") && atoi(getenv("GRAPE_B200_WSEG_SCAN")) == 0);
    // radix-4 levels where five N x N matrices per sub-warp fit the shared memory of a block (N <= 16)
    a.radix = ((size_t)5 * p.N * p.N * sizeof(cplx) * (128 / wp.W) <= 200 * 1024) ? 4 : 2;
    if (const char* e = getenv("GRAPE_B200_WSEG_RADIX")) a.radix = atoi(e) == 4 && a.radix == 4 ? 4 : 2;
    // without chains the segment count only costs scan levels: 256 segments (measured on C2: S = 8 0.146 ms, 12: 0.152,
    // 16: 0.159, 24: 0.179 with radix-2 levels, profiles/r2_s24_bench_c2_*.json)
    if (a.scan) S = std::max(2, (p.NT + 255) / 256);
    if (const char* e = getenv("GRAPE_B200_SEG_S")) S = atoi(e);
    a.S = S < 2 ? 2 : (S > 128 ? 128 : S);
    a.NSEG = (p.NT + a.S - 1) / a.S;
    const size_t pbytes = (size_t)a.NSEG * p.G * p.N * p.N * sizeof(cplx);
    void* q = nullptr;
    if (cudaMalloc(&q, pbytes) != cudaSuccess) { err = "cudaMalloc failed (warp seg)"; return GRAPE_B200_ECUDA; }
    allocs.push_back(q);
    a.Pseg = static_cast<cplx*>(q);
    a.Qa = a.Qb = nullptr;
    a.Q = a.Pseg;
    if (a.scan) {
        for (cplx** dst : {&a.Qa, &a.Qb}) {
            if (cudaMalloc(&q, pbytes) != cudaSuccess) { err = "cudaMalloc failed (warp seg scan)"; return GRAPE_B200_ECUDA; }
            allocs.push_back(q);
            *dst = static_cast<cplx*>(q);
        }
        int levels = 0;
        for (int off = 1; off < a.NSEG; off *= a.radix) ++levels;
        a.Q = levels == 0 ? a.Pseg : ((levels & 1) ? a.Qa : a.Qb);
    }
    cudaError_t e = cudaSuccess;
    a.pf = ((size_t)4 * p.N * p.N * sizeof(cplx) * (128 / wp.W) <= 160 * 1024) ? 1 : 0;   // double-buffered propagators
    const size_t smem = (size_t)(3 + a.pf) * p.N * p.N * sizeof(cplx) * (128 / wp.W);
    WARP_SWITCH(wp.W, e = cudaFuncSetAttribute(warp_segprod<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
    if (e == cudaSuccess && a.scan) {
        const size_t smem2 = (size_t)(a.radix + 1) * p.N * p.N * sizeof(cplx) * (128 / wp.W);
        WARP_SWITCH(wp.W, e = cudaFuncSetAttribute(warp_scan_level<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2))
    }
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute failed (warp seg): ") + cudaGetErrorString(e); return GRAPE_B200_ECUDA; }
    return 0;
}

inline void warp_seg_run_prod(const WarpSegArgs& a, const WarpPlan& wp, const DevP& p, cudaStream_t st, int64_t& launches) {
    const int spb = 128 / wp.W;
    const long long units = (long long)p.G * a.NSEG;
    const size_t smem = (size_t)(3 + a.pf) * p.N * p.N * sizeof(cplx) * spb;
    WARP_SWITCH(wp.W, warp_segprod<WW><<<(unsigned)((units + spb - 1) / spb), 128, smem, st>>>(p, a, spb, a.pf))
    launches++;
    if (a.scan) {
        const size_t smem2 = (size_t)(a.radix + 1) * p.N * p.N * sizeof(cplx) * spb;
        const cplx* in = a.Pseg;
        cplx* out = a.Qa;
        for (int off = 1; off < a.NSEG; off *= a.radix) {
            WARP_SWITCH(wp.W, warp_scan_level<WW><<<(unsigned)((units + spb - 1) / spb), 128, smem2, st>>>(p, a, in, out, off, a.radix, spb))
            launches++;
            in = out;
            out = out == a.Qa ? a.Qb : a.Qa;
        }
    }
}
inline void warp_seg_run_forward(const WarpSegArgs& a, const WarpPlan& wp, const DevP& p, bool fill, cudaStream_t st, int64_t& launches) {
    const int spb = 128 / wp.W;
    if (a.scan) {
        const long long units = (long long)p.K * a.NSEG;
        WARP_SWITCH(wp.W, (warp_scan_bounds_fwd<WW><<<(unsigned)((units + spb - 1) / spb), 128, 0, st>>>(p, a)))
        launches++;
    } else {
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
    if (a.scan) {
        const long long units = (long long)p.K * a.NSEG;
        WARP_SWITCH(wp.W, (warp_scan_bounds_bwd<WW><<<(unsigned)((units + spb - 1) / spb), 128, 0, st>>>(p, a, chi_host)))
    } else {
        const long long units = p.K;
        WARP_SWITCH(wp.W, (warp_seg_bwd<WW, false><<<(unsigned)((units + spb - 1) / spb), 128, 0, st>>>(p, a, chi_host)))
    }
    {
        const long long units = (long long)p.K * a.NSEG;
        WARP_SWITCH(wp.W, (warp_seg_bwd<WW, true><<<(unsigned)((units + spb - 1) / spb), 128, 0, st>>>(p, a, chi_host)))
    }
    launches += 2;
}

