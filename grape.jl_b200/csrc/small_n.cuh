// small_n.cuh -- N <= 4 path: one thread per (generator, step) / trajectory /
// (trajectory, step); every matrix and state lives in registers.
//
// Device layout (structure-of-arrays, fastest index = generator g or trajectory k,
// so that a warp of consecutive k issues 512-byte coalesced double2 accesses):
//   H0s [c][g]        c = i*N + j   (matrix element H_ij)
//   Hcs [l][c][g]
//   U   [n][c][g]     U_{g,n} = exp(-i H_{g,n} dt_n)
//   psi [n][i][k]     n = 0..NT   (fw_storage, reference src/workspace.jl:215)
//   chi [n][i][k]     n = 1..NT   chi_k entering backward step n
//
// Phases:
//   A  small_form_U    : all (g,n) in parallel        -- replaces the exp inside prop_step! (optimize.jl:732)
//   B1 small_forward   : chain over n per trajectory  -- optimize.jl:720-753
//   B2 small_backward  : chi boundary + chain         -- optimize.jl:845-869, 880-881, 897-909
//   C  small_gradient  : all (k,n) in parallel        -- GradGenerator / taylor_grad_step! contraction,
//                                                       optimize.jl:893-895, 604-653, 946-970
#pragma once
#include "common.cuh"

template <int N>
struct SmallCfg {
    static constexpr int BS = (N <= 3) ? 4 : 2;   // Paterson-Stockmeyer block size
};

// C = A*B, all N x N row-major in registers
template <int N>
GB_D void sm_mm(cplx (&C)[N * N], const cplx (&A)[N * N], const cplx (&B)[N * N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < N; ++k) cfma(acc, A[i * N + k], B[k * N + j]);
            C[i * N + j] = acc;
        }
}

template <int N>
GB_D double sm_norm1(const cplx (&A)[N * N]) {
    double nrm = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) s += cabs1(A[i * N + j]);
        nrm = fmax(nrm, s);
    }
    return nrm;
}

// X = exp(A) by scaling-and-squaring Taylor, Paterson-Stockmeyer evaluation.
// On entry P[0] = A (already divided by 2^s). P is scratch.
template <int N, int BS>
GB_D void sm_expm(cplx (&P)[BS][N * N], cplx (&X)[N * N], int degree, int s) {
    constexpr int NN = N * N;
    const int q = (degree + 1) / BS - 1;
#pragma unroll
    for (int t = 1; t < BS; ++t) {
        if (t < BS - 1 || q > 0) {
            if (t == 3) sm_mm<N>(P[t], P[1], P[1]);          // A^4 = A^2 * A^2
            else sm_mm<N>(P[t], P[t - 1], P[0]);
        }
    }
    // X = B_q
    {
        const double c0 = c_invfact[BS * q];
#pragma unroll
        for (int c = 0; c < NN; ++c) X[c] = mk(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < N; ++i) X[i * N + i].x = c0;
#pragma unroll
        for (int t = 1; t < BS; ++t) {
            const double ct = c_invfact[BS * q + t];
#pragma unroll
            for (int c = 0; c < NN; ++c) cfmar(X[c], ct, P[t - 1][c]);
        }
    }
    for (int r = q - 1; r >= 0; --r) {
        double cf[BS];
#pragma unroll
        for (int t = 0; t < BS; ++t) cf[t] = c_invfact[BS * r + t];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            cplx tmp[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx acc = mk(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) cfma(acc, P[BS - 1][i * N + k], X[k * N + j]);
                tmp[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                cplx v = tmp[i];
                if (i == j) v.x += cf[0];
#pragma unroll
                for (int t = 1; t < BS; ++t) cfmar(v, cf[t], P[t - 1][i * N + j]);
                X[i * N + j] = v;
            }
        }
    }
    for (int t = 0; t < s; ++t) {
        sm_mm<N>(P[0], X, X);
#pragma unroll
        for (int c = 0; c < NN; ++c) X[c] = P[0][c];
    }
}

// ---------------------------------------------------------------------------
// Phase A: U_{g,n} = exp(-i (H0_g + sum_l S_ln eps_ln Hc_gl) dt_n)
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) small_form_U(DevP p) {
    constexpr int NN = N * N, BS = SmallCfg<N>::BS;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)p.G * p.NT) return;
    const int G = p.G, NT = p.NT;
    const int g = (int)(idx % G), n = (int)(idx / G);
    const double dt = p.tlist[n + 1] - p.tlist[n];
    cplx P[BS][NN];
    cplx X[NN];
#pragma unroll
    for (int c = 0; c < NN; ++c) X[c] = __ldg(&p.H0[(size_t)c * G + g]);
    for (int l = 0; l < p.L; ++l) {
        double a = p.eps[l * NT + n];
        if (p.shape) a *= p.shape[l * NT + n];
#pragma unroll
        for (int c = 0; c < NN; ++c) cfmar(X[c], a, __ldg(&p.Hc[((size_t)l * NN + c) * G + g]));
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
    cplx* Uo = p.U + (size_t)n * NN * G + g;
#pragma unroll
    for (int c = 0; c < NN; ++c) Uo[(size_t)c * G] = X[c];
}

// g_b = Re <psi|D|psi>,  D at Ds[c*nD + d]
template <int N>
GB_D double sm_quadform(const cplx* __restrict__ Dk, int nD, const cplx (&psi)[N]) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        cplx t = mk(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < N; ++j) cfma(t, __ldg(&Dk[(size_t)(i * N + j) * nD]), psi[j]);
        acc += psi[i].x * t.x + psi[i].y * t.y;
    }
    return acc;
}
// x += f * (-D psi)
template <int N>
GB_D void sm_add_xi(const cplx* __restrict__ Dk, int nD, double f, const cplx (&psi)[N], cplx (&x)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        cplx t = mk(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < N; ++j) cfma(t, __ldg(&Dk[(size_t)(i * N + j) * nD]), psi[j]);
        x[i].x = fma(-f, t.x, x[i].x);
        x[i].y = fma(-f, t.y, x[i].y);
    }
}

// ---------------------------------------------------------------------------
// Phase B1: forward chain. One thread per trajectory; U_n streamed through a
// per-thread cp.async ring in shared memory (depth D), states written with
// coalesced streaming double2 stores.
// ---------------------------------------------------------------------------
template <int N, int D>
__global__ void small_forward(DevP p) {
    constexpr int NN = N * N;
    extern __shared__ __align__(16) unsigned char smraw[];
    cplx* sm = reinterpret_cast<cplx*>(smraw);
    const int BD = blockDim.x, tid = threadIdx.x;
    const int K = p.K, G = p.G, NT = p.NT;
    const int k = blockIdx.x * BD + tid;
    const bool act = k < K;
    const int kk = act ? k : K - 1;
    const int g = p.gen[kk];
    const cplx* Ug = p.U + g;
    auto issue = [&](int n) {
        if (n < NT) {
            const cplx* src = Ug + (size_t)n * NN * G;
            cplx* dst = sm + (size_t)((n % D) * NN) * BD + tid;
#pragma unroll
            for (int c = 0; c < NN; ++c) cp_async16(dst + c * BD, src + (size_t)c * G);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int n = 0; n < D - 1; ++n) issue(n);

    cplx psi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) psi[i] = p.psi0[(size_t)i * K + kk];
    if (act) {
#pragma unroll
        for (int i = 0; i < N; ++i) st_cs(&p.psi[(size_t)i * K + k], psi[i]);
    }
    const bool gb = p.gb_kind != 0;
    const int nD = p.gb_nD;
    const cplx* Dk = gb ? p.D + (nD == 1 ? 0 : kk) : nullptr;
    double jb = 0.0;
    if (gb) jb = sm_quadform<N>(Dk, nD, psi) * ((p.tlist[1] - p.tlist[0]) * 0.5);

    for (int n = 0; n < NT; ++n) {
        issue(n + D - 1);
        cp_async_wait<D - 1>();
        const cplx* u = sm + (size_t)((n % D) * NN) * BD + tid;
        cplx nw[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j) cfma(acc, u[(i * N + j) * BD], psi[j]);
            nw[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) psi[i] = nw[i];
        if (act) {
            cplx* o = p.psi + ((size_t)(n + 1) * N) * K + k;
#pragma unroll
            for (int i = 0; i < N; ++i) st_cs(o + (size_t)i * K, psi[i]);
        }
        if (gb) {
            const int ntl = n + 1;
            const double w = (ntl < NT) ? 0.5 * (p.tlist[ntl + 1] - p.tlist[ntl - 1])
                                        : 0.5 * (p.tlist[NT] - p.tlist[NT - 1]);
            jb = fma(sm_quadform<N>(Dk, nD, psi), w, jb);
        }
    }
    if (act) {
        cplx acc = mk(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < N; ++i) cfmac(acc, p.tgt[(size_t)i * K + k], psi[i]);
        p.tau[k] = acc;
        p.jb[k] = jb;
    }
}

// ---------------------------------------------------------------------------
// Phase B2: chi boundary condition + backward chain chi <- U_n^dagger chi.
// ---------------------------------------------------------------------------
template <int N, int D>
__global__ void small_backward(DevP p, const cplx* __restrict__ chi_host) {
    constexpr int NN = N * N;
    extern __shared__ __align__(16) unsigned char smraw[];
    cplx* sm = reinterpret_cast<cplx*>(smraw);
    const int BD = blockDim.x, tid = threadIdx.x;
    const int K = p.K, G = p.G, NT = p.NT;
    const int k = blockIdx.x * BD + tid;
    const bool act = k < K;
    const int kk = act ? k : K - 1;
    const int g = p.gen[kk];
    const cplx* Ug = p.U + g;
    // ring slot for step n (descending): index by (NT-1-n)
    auto issue = [&](int q) {   // q-th step in backward order, n = NT-1-q
        if (q < NT) {
            const int n = NT - 1 - q;
            const cplx* src = Ug + (size_t)n * NN * G;
            cplx* dst = sm + (size_t)((q % D) * NN) * BD + tid;
#pragma unroll
            for (int c = 0; c < NN; ++c) cp_async16(dst + c * BD, src + (size_t)c * G);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int q = 0; q < D - 1; ++q) issue(q);

    const bool gb = p.gb_kind != 0 && p.lambda_b != 0.0;
    const int nD = p.gb_nD;
    const cplx* Dk = gb ? p.D + (nD == 1 ? 0 : kk) : nullptr;

    // boundary condition chi_k(T)  (optimize.jl:845-869)
    cplx x[N];
    if (chi_host) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = chi_host[(size_t)kk * N + i];
    } else {
        const double w = p.w ? p.w[kk] : 1.0;
        const double Kg = (double)p.Kglobal;
        cplx c;
        if (p.functional == 0) c = mk(w * p.sums[0] / (Kg * Kg), w * p.sums[1] / (Kg * Kg));
        else if (p.functional == 1) c = mk(w / (2.0 * Kg), 0.0);
        else { cplx t = p.tau[kk]; c = mk(w * t.x / Kg, w * t.y / Kg); }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = cmul(c, p.tgt[(size_t)i * K + kk]);
    }
    if (gb) {
        cplx pT[N];
#pragma unroll
        for (int i = 0; i < N; ++i) pT[i] = p.psi[((size_t)NT * N + i) * K + kk];
        sm_add_xi<N>(Dk, nD, p.lambda_b * (p.tlist[NT] - p.tlist[NT - 1]) * 0.5, pT, x);
    }
    double rho = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) rho += cnorm2(x[i]);
    rho = sqrt(rho);
    if (!(rho >= p.chi_min_norm)) {
        if (act && atomicCAS(&p.flags->chi_bad_k, 0, k + 1) == 0) p.flags->chi_bad_rho = rho;
        rho = 1.0;
    }
    {
        const double ir = 1.0 / rho;
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = cscale(x[i], ir);
    }
    if (act) {
        p.rho[k] = rho;
#pragma unroll
        for (int i = 0; i < N; ++i) p.chiT[(size_t)k * N + i] = x[i];
    }

    for (int q = 0; q < NT; ++q) {
        const int n = NT - 1 - q;
        issue(q + D - 1);
        cplx pp[N];
        if (gb && n > 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) pp[i] = ld_cs(&p.psi[((size_t)n * N + i) * K + kk]);
        }
        if (act) {
            cplx* o = p.chi + ((size_t)(n + 1) * N) * K + k;
#pragma unroll
            for (int i = 0; i < N; ++i) st_cs(o + (size_t)i * K, x[i]);
        }
        cp_async_wait<D - 1>();
        const cplx* u = sm + (size_t)((q % D) * NN) * BD + tid;
        cplx nw[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            cplx acc = mk(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < N; ++j) cfmac(acc, u[(j * N + i) * BD], x[j]);   // (U^dagger x)_i
            nw[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = nw[i];
        if (gb && n > 0)   // optimize.jl:897-908: chi += lambda_b * 0.5(t_{n+1}-t_{n-1}) / rho * xi(Psi(t_{n-1}))
            sm_add_xi<N>(Dk, nD, p.lambda_b * 0.5 * (p.tlist[n + 1] - p.tlist[n - 1]) / rho, pp, x);
    }
}

// ---------------------------------------------------------------------------
// Phase C: per (k,n)   w_l = (dU_n^dagger / d eps_l) chi_k(t_n)   by the block
// (GradGenerator) Taylor recursion in vector form, then
//   tau_grad[k][n,l] = rho_k <w_l | Psi_k(t_{n-1})>   and the block-level sum over k.
// Block = BK (trajectories) x 128/BK (steps). Controls [l0, l0+LC).
// ---------------------------------------------------------------------------
template <int N, int LC>
__global__ void __launch_bounds__(128) small_gradient(DevP p, int l0, int BK) {
    constexpr int NN = N * N;
    __shared__ double s_red[4][LC];
    const int tid = threadIdx.x;
    const int tk = tid % BK, tn = tid / BK, BN = 128 / BK;
    const int K = p.K, G = p.G, NT = p.NT;
    const int k = blockIdx.x * BK + tk;
    const int n = blockIdx.y * BN + tn;
    const bool act = (k < K) && (n < NT);
    const int kk = k < K ? k : K - 1;
    const int nn = n < NT ? n : NT - 1;
    const int g = p.gen[kk];
    const double dt = p.tlist[nn + 1] - p.tlist[nn];

    // Abar = +i dt H^dagger : Abar_ij = i dt conj(H_ji)
    cplx A[NN];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) A[i * N + j] = __ldg(&p.H0[(size_t)(j * N + i) * G + g]);
    for (int l = 0; l < p.L; ++l) {
        double a = p.eps[l * NT + nn];
        if (p.shape) a *= p.shape[l * NT + nn];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j)
                cfmar(A[i * N + j], a, __ldg(&p.Hc[((size_t)l * NN + j * N + i) * G + g]));
    }
    int m, s;
    {
        double nrm = dt * sm_norm1<N>(A);
        if (p.grad_method == 0) vec_plan(nrm, m, s);
        else { s = 0; m = p.taylor_max_order; }
    }
    const double sc = dt * ldexp(1.0, -s);
#pragma unroll
    for (int c = 0; c < NN; ++c) A[c] = mk(sc * A[c].y, sc * A[c].x);
    cplx E[LC][NN];
#pragma unroll
    for (int l = 0; l < LC; ++l) {
        double sl = sc;
        if (p.dshape) sl *= p.dshape[(l0 + l) * NT + nn];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                cplx e = __ldg(&p.Hc[((size_t)(l0 + l) * NN + j * N + i) * G + g]);
                E[l][i * N + j] = mk(sl * e.y, sl * e.x);
            }
    }

    cplx asum[N], bsum[LC][N];
#pragma unroll
    for (int i = 0; i < N; ++i) asum[i] = ld_cs(&p.chi[((size_t)(nn + 1) * N + i) * K + kk]);
#pragma unroll
    for (int l = 0; l < LC; ++l)
#pragma unroll
        for (int i = 0; i < N; ++i) bsum[l][i] = mk(0.0, 0.0);

    const bool taylor = p.grad_method != 0;
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
            {
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

    const double rho = p.rho[kk];
    double red[LC];
#pragma unroll
    for (int l = 0; l < LC; ++l) {
        cplx acc = mk(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < N; ++i)
            cfmac(acc, bsum[l][i], ld_cs(&p.psi[((size_t)nn * N + i) * K + kk]));
        acc = cscale(acc, rho);
        if (act && p.taugrads) p.taugrads[((size_t)k * p.L + (l0 + l)) * NT + n] = acc;
        red[l] = act ? acc.x : 0.0;
    }
    // fixed-order reduction over the BK trajectories of this block
    const int width = BK < 32 ? BK : 32;
#pragma unroll
    for (int l = 0; l < LC; ++l)
        for (int off = width >> 1; off > 0; off >>= 1)
            red[l] += __shfl_down_sync(0xffffffffu, red[l], off, width);
    if (BK > 32) {
        const int warp = tid >> 5;
        if ((tid & 31) == 0)
#pragma unroll
            for (int l = 0; l < LC; ++l) s_red[warp][l] = red[l];
        __syncthreads();
        if (tk == 0 && n < NT) {
            const int nw = BK >> 5;
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                double sacc = 0.0;
                for (int w = 0; w < nw; ++w) sacc += s_red[tn * nw + w][l];
                p.partial[(size_t)blockIdx.x * p.L * NT + (size_t)(l0 + l) * NT + n] = sacc;
            }
        }
    } else if (tk == 0 && n < NT) {
#pragma unroll
        for (int l = 0; l < LC; ++l)
            p.partial[(size_t)blockIdx.x * p.L * NT + (size_t)(l0 + l) * NT + n] = red[l];
    }
}
