// dense_kry.cuh -- Krylov-form gradient contraction of the large-N path.
//
// The reference propagates, per trajectory and time step, the GradGenerator block vector
// [chi'_1 .. chi'_L, chi] with a dense exponential of the N(L+1) block matrix (src/optimize.jl:880-911,
// docs/src/background.md:447-494); the block-recursion kernels (dense.cuh / dense2.cuh) do the same with
// 1 + 2L operator applications per Taylor order.  With A = +i dt H^dagger and E_l = +i dt s_l mu_l^dagger the
// block action is chi'_l = L_exp(A)[E_l] chi (Frechet derivative), and
//     <chi'_l | Psi> = sum_{a,b} a! b!/(a+b+1)!  ch_b^dagger E_l^dagger bh_a = Tr(E_l^dagger M),
//     bh_a = (-i H dt)^a Psi / a!,  ch_b = (+i H^dagger dt)^b chi / b!,  M = sum_{a,b} beta(a,b) bh_a ch_b^dagger,
// where bh_a are exactly the Taylor terms of the forward sweep and ch_b those of a backward sweep of chi alone.
// So, independently of the number of controls:
//   1. both chains run on K columns (dense_chain / dense2_chain) and leave their terms in HBM slots,
//   2. kry_combine   e_b = rho_k sum_{a <= m-1-b} beta(a,b) bh_a            (element-wise, all steps in parallel)
//   3. kry_contract  M_n = sum_{b,k} e_{b,k} ch_{b,k}^dagger  as a 64x64-tiled FP64 DMMA GEMM over the
//                    contraction index (b,k) for ALL time steps in parallel (no grid barrier), with the epilogue
//                    g[n,l] = sum_pq Re(conj(E_l[p,q]) M_n[p,q]) against the control operators,
//   4. kry_reduce    fixed-order sum over the tiles -> partial[l][n]  (then finalize_grad, optimize.jl:574-584).
// The truncation a + b <= m - 1 is the one of the m-term block recursion, so both forms agree to rounding.
// Steps that need sub-stepping (||H dt|| > 1) or more than MT orders, and gradient_method = :taylor, use the
// block recursion: kry_plan decides on the device, per call, which set of kernels does the work.
#pragma once
#include "dense.cuh"
#include "dense2.cuh"

constexpr int KM_T = 64;     // contraction tile: 64 x 64 elements of M per CTA pass
constexpr int KM_ST = 4;     // cp.async stages
constexpr int KM_THREADS = 256;

// beta(a,b) = a! b! / (a+b+1)!
struct KryBeta {
    double v[KRY_MTMAX][KRY_MTMAX];
    constexpr KryBeta() : v() {
        for (int a = 0; a < KRY_MTMAX; ++a)
            for (int b = 0; b < KRY_MTMAX; ++b) {
                double c = 1.0;   // C(a+b, a)
                for (int t = 1; t <= a; ++t) c = c * (double)(b + t) / (double)t;
                v[a][b] = 1.0 / ((double)(a + b + 1) * c);
            }
    }
};
__constant__ KryBeta c_kbeta = KryBeta();

// Taylor order of every step and the per-call decision (single block)
__global__ void __launch_bounds__(256) kry_plan(DevP p, DenseDev d, KryDev kd) {
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    for (int n = threadIdx.x; n < p.NT; n += blockDim.x) {
        int m, s;
        const double* gw;
        dense_plan(p, d, n, p.tlist[n + 1] - p.tlist[n], m, s, gw, d.econ != 0);
        kd.m_n[n] = m;
        if (s != 0 || m > kd.MT) atomicOr(&s_bad, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ok = s_bad ? 0 : 1;
        *kd.ok = ok;
        *kd.kb = ok ? 1 : p.KB;
    }
}

// e_b = rho_k sum_{a=0}^{m-1-b} beta(a,b) bh_a, in place in the forward term slots (bh_0 = fw_storage[n]).
// Concurrent chains (kd.conc, dense_chain<2, NS>): the chi chain carried tgt_k / ||tgt_k||, and the complex factor
// f_k = ||tgt_k|| conj(c_k) replaces rho_k (M_n = sum e_b ch_b^dagger is anti-linear in chi).
__global__ void __launch_bounds__(256) kry_combine(DevP p, DenseDev d, KryDev kd, int conc) {
    if (!(*kd.ok)) return;
    const size_t splane = (size_t)d.Np * d.Kp, slot = 2 * splane;
    if (!conc) {
        const size_t total = (size_t)p.NT * slot;
        for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
            const int n = (int)(e / slot);
            const size_t r = e % slot;
            const int k = (int)(r % d.Kp);
            const int m = kd.m_n[n];
            const double* gw = d.econ ? c_econ.g[m] : c_econ.ones;   // a call served here has s == 0 in every step: all economised
            const double rho = k < p.K ? p.rho[k] : 0.0;
            double* ft = kd.FT + (size_t)n * kd.MT * slot + r;
            double v[KRY_MTMAX];
            v[0] = __ldcs(&d.store[(size_t)n * slot + r]);
#pragma unroll
            for (int a = 1; a < KRY_MTMAX; ++a) v[a] = a < m ? __ldcs(&ft[(size_t)a * slot]) : 0.0;
#pragma unroll
            for (int b = 0; b < KRY_MTMAX; ++b) {
                if (b < m) {
                    double s = 0.0;
#pragma unroll
                    for (int a = 0; a < KRY_MTMAX - b; ++a)
                        if (a < m - b) s = fma(c_kbeta.v[a][b] * gw[a + b + 1], v[a], s);
                    ft[(size_t)b * slot] = rho * s;
                }
            }
        }
        return;
    }
    const size_t total = (size_t)p.NT * splane;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(e / splane);
        const size_t r = e % splane;
        const int k = (int)(r % d.Kp);
        const int m = kd.m_n[n];
        const double* gw = d.econ ? c_econ.g[m] : c_econ.ones;
        const double fr = k < p.K ? kd.kfac[k] : 0.0, fi = k < p.K ? kd.kfac[d.Kp + k] : 0.0;
        double* ft = kd.FT + (size_t)n * kd.MT * slot + r;
        double vr[KRY_MTMAX], vi[KRY_MTMAX];
        vr[0] = __ldcs(&d.store[(size_t)n * slot + r]);
        vi[0] = __ldcs(&d.store[(size_t)n * slot + splane + r]);
#pragma unroll
        for (int a = 1; a < KRY_MTMAX; ++a) {
            vr[a] = a < m ? __ldcs(&ft[(size_t)a * slot]) : 0.0;
            vi[a] = a < m ? __ldcs(&ft[(size_t)a * slot + splane]) : 0.0;
        }
#pragma unroll
        for (int b = 0; b < KRY_MTMAX; ++b) {
            if (b < m) {
                double sr = 0.0, si = 0.0;
#pragma unroll
                for (int a = 0; a < KRY_MTMAX - b; ++a)
                    if (a < m - b) {
                        const double bg = c_kbeta.v[a][b] * gw[a + b + 1];
                        sr = fma(bg, vr[a], sr);
                        si = fma(bg, vi[a], si);
                    }
                ft[(size_t)b * slot] = fr * sr - fi * si;
                ft[(size_t)b * slot + splane] = fr * si + fi * sr;
            }
        }
    }
}

struct KAcc { double re[2], im[2]; };

// g[n,l] tile partials.  One 64x64 tile of M_n = X_n Y_n^dagger per pass, X = e (FT slots), Y = ch (BT slots),
// contraction over (b, k): slot b, columns k of the planar [Np][Kp] term layout (k contiguous), so both DMMA
// operands are read "row-major with the contraction index fastest" straight from the chains' output.
template <int KC>
__global__ void __launch_bounds__(KM_THREADS, 1) kry_contract(DevP p, DenseDev d, KryDev kd) {
    if (!(*kd.ok)) return;
    extern __shared__ __align__(16) double ksm[];
    __shared__ double s_buf[32 * DENSE_LMAX];
    constexpr int AS = KC + 4;               // padded row stride: conflict-free 64-bit fragment loads
    constexpr int PLANE = KM_T * AS;         // one plane (re or im) of one operand tile
    constexpr int STAGE = 4 * PLANE;         // X re, X im, Y re, Y im
    constexpr int SPR = KC / 2;              // 16-byte segments per tile row
    constexpr int SEGS = 4 * KM_T * SPR;
    const int Np = d.Np, Kp = d.Kp, NT = p.NT, L = p.L, MT = kd.MT, TP = kd.TP, TT = kd.TT;
    const size_t splane = (size_t)Np * Kp, slot = 2 * splane, hplane = (size_t)Np * Np;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3, wr = w >> 2, wc = w & 3;   // warp tile: rows wr*32.., cols wc*16..
    const int cpk = Kp / KC;
    const long long total = (long long)NT * TT;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int n = (int)(t / TT), r = (int)(t % TT);
        const int p0 = (r / TP) * KM_T, q0 = (r % TP) * KM_T;
        const int nst = kd.m_n[n] * cpk;
        const double* Xn = kd.FT + (size_t)n * MT * slot;
        const double* Yn = kd.BT + (size_t)n * MT * slot;
        auto issue = [&](int sidx) {
            const int b = sidx / cpk, k0 = (sidx % cpk) * KC;
            double* st = ksm + (size_t)(sidx % KM_ST) * STAGE;
#pragma unroll
            for (int e0 = 0; e0 < SEGS; e0 += KM_THREADS) {
                const int e = e0 + threadIdx.x;
                const int seg = e % SPR, row = (e / SPR) % KM_T, pl = (e / (SPR * KM_T)) & 1, op = e / (SPR * KM_T * 2);
                const int grow = (op ? q0 : p0) + row;
                const bool valid = grow < Np;
                const double* src = (op ? Yn : Xn) + (size_t)b * slot + (size_t)pl * splane +
                                    (size_t)(valid ? grow : 0) * Kp + k0 + 2 * seg;
                cp_async16_zfill(st + (op * 2 + pl) * PLANE + row * AS + 2 * seg, src, valid);
            }
        };
        KAcc acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j].re[0] = acc[i][j].re[1] = acc[i][j].im[0] = acc[i][j].im[1] = 0.0;
#pragma unroll
        for (int s = 0; s < KM_ST - 1; ++s) {
            if (s < nst) issue(s);
            cp_async_commit();
        }
        for (int c = 0; c < nst; ++c) {
            cp_async_wait<KM_ST - 2>();
            __syncthreads();
            const int nx = c + KM_ST - 1;
            if (nx < nst) issue(nx);
            cp_async_commit();
            const double* st = ksm + (size_t)(c % KM_ST) * STAGE;
#pragma unroll
            for (int kk = 0; kk < KC / 4; ++kk) {
                double xr[4], xi[4], yr[2], yi[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int o = (wr * 32 + i * 8 + lr) * AS + kk * 4 + lc;
                    xr[i] = st[o];
                    xi[i] = st[PLANE + o];
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int o = (wc * 16 + j * 8 + lr) * AS + kk * 4 + lc;
                    yr[j] = st[2 * PLANE + o];
                    yi[j] = st[3 * PLANE + o];
                }
                // M = X conj(Y)^T:  M_re += xr yr + xi yi,  M_im += xi yr - xr yi
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        dmma884(acc[i][j].re, xr[i], yr[j]);
                        dmma884(acc[i][j].im, xi[i], yr[j]);
                    }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double nxr = -xr[i];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        dmma884(acc[i][j].re, xi[i], yi[j]);
                        dmma884(acc[i][j].im, nxr, yi[j]);
                    }
                }
            }
        }
        cp_async_wait<0>();
        // epilogue: Re(conj(E_l) M) with E_l = i dt s_l a, a = mu_l^dagger:  a_re M_im - a_im M_re  (dt s_l in kry_reduce)
        double gv[DENSE_LMAX];
#pragma unroll
        for (int l = 0; l < DENSE_LMAX; ++l) gv[l] = 0.0;
#pragma unroll
        for (int l = 0; l < DENSE_LMAX; ++l) {
            if (l < L) {
                const double* Ar = d.Ha + (size_t)(1 + l) * 2 * hplane;
                const double* Ai = Ar + hplane;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int prow = p0 + wr * 32 + i * 8 + lr, qcol = q0 + wc * 16 + j * 8 + 2 * lc;
                        if (prow < Np && qcol < Np) {
                            const double2 ar = __ldg(reinterpret_cast<const double2*>(&Ar[(size_t)prow * Np + qcol]));
                            const double2 ai = __ldg(reinterpret_cast<const double2*>(&Ai[(size_t)prow * Np + qcol]));
                            gv[l] = fma(ar.x, acc[i][j].im[0], gv[l]);
                            gv[l] = fma(-ai.x, acc[i][j].re[0], gv[l]);
                            gv[l] = fma(ar.y, acc[i][j].im[1], gv[l]);
                            gv[l] = fma(-ai.y, acc[i][j].re[1], gv[l]);
                        }
                    }
            }
        }
        block_sum<DENSE_LMAX>(gv, s_buf);   // ends with __syncthreads: the stage buffers are free again
        if (threadIdx.x == 0) {
#pragma unroll
            for (int l = 0; l < DENSE_LMAX; ++l)
                if (l < L) kd.tilepart[((size_t)n * TT + r) * L + l] = gv[l];
        }
    }
}

// partial[l][n] = dt_n s_{l,n} sum_tiles  (fixed order); finalize_grad applies -2 (optimize.jl:574-584)
__global__ void __launch_bounds__(256) kry_reduce(DevP p, DenseDev d, KryDev kd) {
    if (!(*kd.ok)) return;
    const int LNT = p.L * p.NT;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < LNT; idx += gridDim.x * blockDim.x) {
        const int l = idx / p.NT, n = idx % p.NT;
        double s = 0.0;
        for (int r = 0; r < kd.TT; ++r) s += kd.tilepart[((size_t)n * kd.TT + r) * p.L + l];
        const double sl = p.dshape ? p.dshape[idx] : 1.0;
        p.partial[idx] = (p.tlist[n + 1] - p.tlist[n]) * sl * s;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// Decides whether this handle gets the Krylov-form backward and allocates the term storage.
// GRAPE_B200_KRYLOV=0 disables it; GRAPE_B200_KRY_MT sets the number of term slots per step.
inline int kry_setup(DensePlan& dp, Dense2Plan& q, DevP& p, std::vector<void*>& allocs, std::string& err) {
    KryDev& kd = dp.kd;
    DenseDev& d = dp.d;
    memset(&kd, 0, sizeof kd);
    d.kry_ok = nullptr;
    const char* env = getenv("GRAPE_B200_KRYLOV");
    if (env && atoi(env) == 0) return 0;
    if (p.grad_method != 0) return 0;               // :taylor reproduces taylor_grad_step! with the block recursion
    if (!q.on && !dp.strip_ok) return 0;
    if (q.on && q.d2.ntiles > q.grid) { /* several tiles per CTA: still fine, the chain keeps no per-tile registers */ }
    int MT = 18;
    if (const char* e = getenv("GRAPE_B200_KRY_MT")) MT = atoi(e);
    MT = std::max(2, std::min(MT, KRY_MTMAX));
    const size_t slot = 2 * (size_t)d.Np * d.Kp;
    size_t freeB = 0, totB = 0;
    if (cudaMemGetInfo(&freeB, &totB) != cudaSuccess) { cudaGetLastError(); return 0; }
    // both term stores must fit in 70 % of what is free now; otherwise fewer slots, then the block recursion
    while (MT >= 12 && 2.0 * (double)p.NT * MT * slot * sizeof(double) > 0.7 * (double)freeB) MT -= 2;
    if (2.0 * (double)p.NT * MT * slot * sizeof(double) > 0.7 * (double)freeB) return 0;
    kd.MT = MT;
    kd.TP = (d.Np + KM_T - 1) / KM_T;
    kd.TT = kd.TP * kd.TP;
    kd.KC = (d.Kp % 16 == 0) ? 16 : 8;
    auto ald = [&](void** dst, size_t bytes) -> int {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, bytes) != cudaSuccess) { cudaGetLastError(); err = "cudaMalloc failed (Krylov-form term storage)"; return GRAPE_B200_ECUDA; }
        allocs.push_back(ptr);
        if (cudaMemset(ptr, 0, bytes) != cudaSuccess) { err = "cudaMemset failed"; return GRAPE_B200_ECUDA; }
        *dst = ptr;
        return 0;
    };
    int rc;
    if ((rc = ald((void**)&kd.m_n, sizeof(int) * (size_t)p.NT))) return rc;
    if ((rc = ald((void**)&kd.ok, sizeof(int)))) return rc;
    if ((rc = ald((void**)&kd.kb, sizeof(int)))) return rc;
    if ((rc = ald((void**)&kd.kcur, sizeof(double) * slot))) return rc;
    if ((rc = ald((void**)&kd.kcur2, sizeof(double) * slot))) return rc;
    if ((rc = ald((void**)&kd.tilepart, sizeof(double) * (size_t)p.NT * kd.TT * p.L))) return rc;
    if ((rc = ald((void**)&kd.FT, sizeof(double) * (size_t)p.NT * MT * slot))) return rc;
    if ((rc = ald((void**)&kd.BT, sizeof(double) * (size_t)p.NT * MT * slot))) return rc;
    dp.kry_smem = sizeof(double) * (size_t)KM_ST * 4 * KM_T * (kd.KC + 4);
    cudaError_t e = kd.KC == 16
        ? cudaFuncSetAttribute(kry_contract<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.kry_smem)
        : cudaFuncSetAttribute(kry_contract<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.kry_smem);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute failed (kry_contract): ") + cudaGetErrorString(e); return GRAPE_B200_ECUDA; }
    kd.on = 1;
    d.kry_ok = kd.ok;
    p.KBdev = kd.kb;
    return 0;
}

// before the forward sweep of every call
inline void kry_run_plan(DensePlan& dp, const DevP& p, cudaStream_t st, int64_t& launches) {
    if (!dp.kd.on) return;
    kry_plan<<<1, 256, 0, st>>>(p, dp.d, dp.kd);
    launches++;
}

// after the backward chain: combine, contract, reduce (each returns at once if the call took the block recursion)
inline void kry_run_gradient(DensePlan& dp, const DevP& p, cudaStream_t st, int64_t& launches, bool conc = false) {
    if (!dp.kd.on) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    kry_combine<<<sms * 8, 256, 0, st>>>(p, dp.d, dp.kd, conc ? 1 : 0);
    const long long total = (long long)p.NT * dp.kd.TT;
    const int grid = (int)std::min<long long>(total, sms);
    if (dp.kd.KC == 16) kry_contract<16><<<grid, KM_THREADS, dp.kry_smem, st>>>(p, dp.d, dp.kd);
    else kry_contract<8><<<grid, KM_THREADS, dp.kry_smem, st>>>(p, dp.d, dp.kd);
    const int blocks = (p.L * p.NT + 255) / 256;
    kry_reduce<<<blocks < 296 ? blocks : 296, 256, 0, st>>>(p, dp.d, dp.kd);
    launches += 3;
}
