// capi.cu -- C-ABI (include/grape_b200.h) of the B200-native GRAPE gradient engine.
//
// Handle = device-resident GrapeWrk (reference src/workspace.jl:78-144): pulse
// values, propagators, forward storage, chi states, gradient buffers.  One
// eval_fg call = one H2D copy of the pulse values, a fixed sequence of kernels
// on one stream, one D2H copy of (G, grad_J_Tb, grad_J_a, J_parts, sums, tau,
// flags) and one stream synchronisation.
#include "../../include/grape_b200.h"
#include "common.cuh"
#include "reduce.cuh"
#include "xchg.cuh"
#include "small_n.cuh"
#include "small_seg.cuh"
#include "small_sym.cuh"
#include "warp_n.cuh"
#include "warp_seg.cuh"
#include "dense.cuh"
#include "dense2.cuh"
#include "dense_kry.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_create_error;

struct grape_b200_handle_impl {
    DevP p;
    int path, device;
    int LNT;
    cudaStream_t stream;
    std::string err;
    std::vector<void*> dev_allocs;
    // contiguous device output block: grad[3*LNT] | Jparts[3] | sums[4] | (pad) | tau[2K] | flags
    double* d_out;
    size_t out_doubles, off_J, off_sums, off_tau, off_flags;
    double* h_out;      // pinned mirror of d_out
    double* h_in;       // pinned staging: pulsevals[LNT] | sums[4]
    double* d_eps_own;  // handle-owned pulse buffer
    double* d_damp;     // handle-owned [L*NT] d a / d eps of amplitude mode (allocated on first use)
    cplx* d_chi_host;   // [K*N] user chi for backward_chi
    cplx* d_tmp;        // gather scratch
    size_t tmp_elems;
    cudaEvent_t ev[8];
    bool profiling, forward_done, backward_done;
    double timings[8];
    int64_t launches;
    WarpPlan warp;
    DensePlan dense;
    Dense2Plan dense2;
    // CUDA graphs of the complete eval_f / eval_fg call (H2D, kernel sequence, D2H): the small and
    // sub-warp paths launch ~10 short kernels per call and are launch-bound for few trajectories
    cudaGraphExec_t graph_fg, graph_f;
    int64_t graph_fg_launches, graph_f_launches;
    bool graphs_ok;
    bool wseg_on;         // warp path: time-segmented schedule (warp_seg.cuh)
    WarpSegArgs wseg;
    bool seg_on;          // small path: time-segmented schedule (small_seg.cuh)
    bool interior_done;   // small path, segmented: fw_storage filled inside the segments
    bool seg_herm;        // small path, segmented, all generators Hermitian (N <= 3): the gradient kernel recomputes
                          // the forward states backwards; fw_storage is only filled when somebody reads it
    bool U_valid;         // small path: p.U holds the propagators of the current pulses
    bool seg_fuse;        // small path, segmented: fused propagator formation + segment product (small_formseg)
    bool seg_scan;        // small path, real-symmetric generators: prefix products by a parallel scan, no boundary chains
    bool bounds_done;     // scan schedule: the segment boundaries of fw_storage hold the states of the current pulses
    bool defer_tau;       // this gradient call forms the tau sums in the finalize kernel (uncoupled functional, one device)
    bool chain_dual;      // chain schedule: the forward phase of this call carried the targets backwards too (small_segchain_dual)
    bool chib_done;       // scan schedule: chiE / rho / chiT were written by small_scan_bounds for the current backward call
    int sym_v;            // 2: operators staged in shared memory (small_*_sym2, default); 1: round-1 kernels (GRAPE_B200_SYM_V=1)
    int sym_occ;          // resident CTAs per SM small_seggrad_sym is compiled for (3; GRAPE_B200_SYM_OCC=2: no spills, 8 warps)
    bool seg_real;        // seg_herm and every generator real (symmetric): real-arithmetic kernels of small_sym.cuh
    cplx* d_taugrads;     // [K][L][NT] dump buffer of get_tau_grads (allocated on first use)
    bool taugrads_valid;  // d_taugrads holds the tau_grads of the last backward sweep
    SegArgs seg;
    // peer exchange over NVLink (xchg.cuh): the shards of a trajectory-sharded problem reduce sums / gradient themselves
    XchgDev xd;
    bool xchg_on;         // peers attached (grape_b200_xchg_attach / grape_b200_multi_create)
    bool xchg_mode;       // set by the entry point: true = exchange inside the call (eval_*, enqueue_*), false = local
                          // partial results (split host API forward / backward / backward_chi)
    bool xchg_fonly;      // this call is a functional-only evaluation (the sums are always exchanged)
    void* xchg_buf;       // this handle's exchange buffer: flags | slots | epochs | timeout
    size_t xchg_bytes;
    std::vector<void*> xchg_ipc_opened;   // peer buffers opened with cudaIpcOpenMemHandle
    bool launched_via_graph;              // eval_launch served the call with a graph launch
    int64_t launch_l0;
};
typedef grape_b200_handle_impl H;

#define CUDA_TRY(h, call)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            char _b[512];                                                                   \
            snprintf(_b, sizeof _b, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),      \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                           \
            (h)->err = _b;                                                                  \
            return GRAPE_B200_ECUDA;                                                        \
        }                                                                                   \
    } while (0)

template <typename T>
int dev_alloc(H* h, T** ptr, size_t count) {
    void* q = nullptr;
    CUDA_TRY(h, cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
    h->dev_allocs.push_back(q);
    *ptr = static_cast<T*>(q);
    return 0;
}
template <typename T>
int dev_upload(H* h, T** ptr, const T* src, size_t count) {
    int rc = dev_alloc(h, ptr, count);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpy(*ptr, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

int choose_path(int N, int requested) {
    if (requested == GRAPE_B200_PATH_SMALL_CHAIN) return GRAPE_B200_PATH_SMALL;
    if (requested == GRAPE_B200_PATH_WARP_CHAIN) return GRAPE_B200_PATH_WARP;
    if (requested != GRAPE_B200_PATH_AUTO) return requested;
    if (N <= 4) return GRAPE_B200_PATH_SMALL;
    if (N <= WARP_MAX_N) return GRAPE_B200_PATH_WARP;
    return GRAPE_B200_PATH_DENSE;
}

// ------------------------------------------------------------------ small path
constexpr int SMALL_D = 8;   // cp.async ring depth

int small_bd(const H* h) { return h->p.K <= 16384 ? 32 : 64; }
size_t small_smem(const H* h) {
    return (size_t)SMALL_D * h->p.N * h->p.N * sizeof(cplx) * small_bd(h);
}

template <int N>
int small_set_attrs(H* h) {
    const int smem = (int)small_smem(h);
    CUDA_TRY(h, cudaFuncSetAttribute(small_forward<N, SMALL_D>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(h, cudaFuncSetAttribute(small_backward<N, SMALL_D>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (N <= 3)
        CUDA_TRY(h, cudaFuncSetAttribute(small_segchain_dual<(N <= 3 ? N : 1)>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_ring_bytes(N)));
    return 0;
}

int small_setup(H* h, const grape_b200_problem* d) {
    DevP& p = h->p;
    const int K = p.K, N = p.N, L = p.L, NT = p.NT, G = p.G, NN = N * N;
    std::vector<cplx> buf;
    // H0s[c][g], c = i*N+j ; ABI: H0[g][j*N+i]
    buf.assign((size_t)NN * G, mk(0, 0));
    for (int g = 0; g < G; ++g)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                const double* s = d->H0 + 2 * ((size_t)g * NN + (size_t)j * N + i);
                buf[(size_t)(i * N + j) * G + g] = mk(s[0], s[1]);
            }
    cplx* q;
    if (int rc = dev_upload(h, &q, buf.data(), buf.size())) return rc;
    p.H0 = q;
    buf.assign((size_t)L * NN * G, mk(0, 0));
    for (int g = 0; g < G; ++g)
        for (int l = 0; l < L; ++l)
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) {
                    const double* s = d->Hc + 2 * (((size_t)g * L + l) * NN + (size_t)j * N + i);
                    buf[((size_t)l * NN + i * N + j) * G + g] = mk(s[0], s[1]);
                }
    if (int rc = dev_upload(h, &q, buf.data(), buf.size())) return rc;
    p.Hc = q;
    auto soa_states = [&](const double* src, const cplx** dst) -> int {
        std::vector<cplx> b((size_t)N * K);
        for (int k = 0; k < K; ++k)
            for (int i = 0; i < N; ++i) b[(size_t)i * K + k] = mk(src[2 * ((size_t)k * N + i)], src[2 * ((size_t)k * N + i) + 1]);
        cplx* qq;
        if (int rc = dev_upload(h, &qq, b.data(), b.size())) return rc;
        *dst = qq;
        return 0;
    };
    if (int rc = soa_states(d->psi0, &p.psi0)) return rc;
    if (int rc = soa_states(d->tgt, &p.tgt)) return rc;
    if (p.gb_kind) {
        const int nD = p.gb_nD;
        buf.assign((size_t)NN * nD, mk(0, 0));
        for (int dd = 0; dd < nD; ++dd)
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) {
                    const double* s = d->gb_D + 2 * ((size_t)dd * NN + (size_t)j * N + i);
                    buf[(size_t)(i * N + j) * nD + dd] = mk(s[0], s[1]);
                }
        if (int rc = dev_upload(h, &q, buf.data(), buf.size())) return rc;
        p.D = q;
    }
    if (int rc = dev_alloc(h, &p.U, (size_t)NT * NN * G)) return rc;
    if (int rc = dev_alloc(h, &p.psi, (size_t)(NT + 1) * N * K)) return rc;
    if (d->gb_kind != 0 || d->path == GRAPE_B200_PATH_SMALL_CHAIN)
        if (int rc = dev_alloc(h, &p.chi, (size_t)(NT + 1) * N * K)) return rc;
    int BK = 1;
    while (BK < K && BK < 128) BK <<= 1;
    p.KB = (K + BK - 1) / BK;
    // time-segmented schedule unless a state running cost couples chi to the stored states at
    // every step (optimize.jl:897-908) or the caller forces the plain chains
    h->seg_on = (p.gb_kind == 0) && (d->path != GRAPE_B200_PATH_SMALL_CHAIN);
    if (h->seg_on) {
        SegArgs& a = h->seg;
        a.BKL = 1;
        while (a.BKL < K && a.BKL < 32) a.BKL <<= 1;
        {
            // Hermitian generators (the usual closed-system case): exp(-iH dt) is unitary and the forward state can be
            // carried backwards next to chi. Exact test on the caller's matrices; GRAPE_B200_SEG_HERM=0 disables.
            bool herm = N <= 3 && !(getenv("GRAPE_B200_SEG_HERM") && atoi(getenv("GRAPE_B200_SEG_HERM")) == 0);
            auto check = [&](const double* M, size_t count) {
                for (size_t g = 0; g < count && herm; ++g)
                    for (int i = 0; i < N && herm; ++i)
                        for (int j = 0; j <= i; ++j) {
                            const double* x = M + 2 * (g * NN + (size_t)j * N + i);
                            const double* y = M + 2 * (g * NN + (size_t)i * N + j);
                            const double tol = 1e-15 * (std::fabs(x[0]) + std::fabs(x[1]) + std::fabs(y[0]) + std::fabs(y[1]));
                            if (std::fabs(x[0] - y[0]) > tol || std::fabs(x[1] + y[1]) > tol) { herm = false; break; }
                        }
            };
            check(d->H0, (size_t)G);
            check(d->Hc, (size_t)G * L);
            a.herm = herm ? 1 : 0;
            a.store_U = herm ? 0 : 1;
            h->seg_herm = herm;
            // real symmetric generators: real-arithmetic segment products and gradient contraction (small_sym.cuh).
            // Exact test (every imaginary part is zero); GRAPE_B200_SEG_REAL=0 disables.
            bool real = herm && !(getenv("GRAPE_B200_SEG_REAL") && atoi(getenv("GRAPE_B200_SEG_REAL")) == 0);
            for (size_t e = 0; e < (size_t)G * NN && real; ++e) real = d->H0[2 * e + 1] == 0.0;
            for (size_t e = 0; e < (size_t)G * L * NN && real; ++e) real = d->Hc[2 * e + 1] == 0.0;
            h->seg_real = real;
            h->sym_occ = getenv("GRAPE_B200_SYM_OCC") ? atoi(getenv("GRAPE_B200_SYM_OCC")) : 3;
            h->sym_v = getenv("GRAPE_B200_SYM_V") ? atoi(getenv("GRAPE_B200_SYM_V")) : 2;
            if (getenv("GRAPE_B200_SYM_OCC") && h->sym_occ < 4) h->sym_v = 1;   // 2 / 3: occupancy variants of the round-1 kernels; 4: A/B variant of the staged gradient kernel
            if (real && h->sym_v != 1) {
                // per-thread operator tile of the staged kernels: (1 + L) N^2 doubles x 128 threads
                const size_t sm = sym_stage_bytes(N, L);
                if (sm > 96 * 1024) h->sym_v = 1;
                else if (sm > 48 * 1024) {
                    cudaError_t e = cudaSuccess;
                    auto set = [&](const void* fn) { if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); };
                    switch (N) {
                        case 1: set((const void*)small_formseg_sym2<1, 0>); set((const void*)small_seggrad_sym2<1, 0>); break;
                        case 2: set((const void*)small_formseg_sym2<2, 0>); set((const void*)small_seggrad_sym2<2, 0>); break;
                        default: set((const void*)small_formseg_sym2<3, 0>); set((const void*)small_seggrad_sym2<3, 0>); break;
                    }
                    if (e != cudaSuccess) { cudaGetLastError(); h->sym_v = 1; }
                }
            }
            if (real) {
                // series tables of the real-symmetric kernels: economised polynomial (default) or Taylor (GRAPE_B200_ECON=0);
                // constant memory is per device and process-wide: handles created with different settings must not be mixed
                const char* ee = getenv("GRAPE_B200_ECON");
                if (sym_tables_upload(!(ee && atoi(ee) == 0)) != cudaSuccess) { h->err = "cudaMemcpyToSymbol failed (series tables)"; return GRAPE_B200_ECUDA; }
            }
            if (real) {
                std::vector<double> rb((size_t)NN * G);
                for (int g = 0; g < G; ++g)
                    for (int i = 0; i < N; ++i)
                        for (int j = 0; j < N; ++j)
                            rb[(size_t)(i * N + j) * G + g] = d->H0[2 * ((size_t)g * NN + (size_t)j * N + i)];
                double* rq;
                if (int rc = dev_upload(h, &rq, rb.data(), rb.size())) return rc;
                a.H0r = rq;
                rb.assign((size_t)L * NN * G, 0.0);
                for (int g = 0; g < G; ++g)
                    for (int l = 0; l < L; ++l)
                        for (int i = 0; i < N; ++i)
                            for (int j = 0; j < N; ++j)
                                rb[((size_t)l * NN + i * N + j) * G + g] = d->Hc[2 * (((size_t)g * L + l) * NN + (size_t)j * N + i)];
                if (int rc = dev_upload(h, &rq, rb.data(), rb.size())) return rc;
                a.Hcr = rq;
                if (int rc = dev_alloc(h, &a.notfast, 1)) return rc;
                CUDA_TRY(h, cudaMemset(a.notfast, 0, sizeof(int)));
            }
        }
        int S = (int)std::ceil(std::sqrt((double)NT));
        {
            // GPU-filling ensembles: the two heavy kernels (formseg, seggrad: 255 registers, 8 resident warps per
            // SM; 12 for the real-symmetric kernels) run in whole waves of warps, each wave taking S steps, while the boundary chains take NSEG steps
            // of about a quarter of that cost: pick the segment length that minimises waves*S + NSEG/4.
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
            const long long resident = (long long)sms * (h->seg_real ? 12 : 8);   // small_sym.cuh kernels: 168 registers
            const long long KGR = (K + a.BKL - 1) / a.BKL, SPW = 32 / a.BKL;
            auto cost = [&](int s_, bool& full) {
                const long long nseg = (NT + s_ - 1) / s_;
                const long long warps = KGR * ((nseg + SPW - 1) / SPW);
                full = warps >= resident;
                // warps drain continuously rather than in lock-step waves: fractional waves, plus half a segment of tail
                // (measured on C3, profiles/r2_s1_c3_sweep.txt: S = 20..25 beats the S = 38 of a whole-wave model by 2 %)
                return std::max(1.0, (double)warps / (double)resident) * s_ + 0.5 * s_ + 0.12 * (double)nseg;   // 0.12: one chain segment (two-warp ring kernel)
            };
            bool full = false;
            double best = cost(S, full);
            if (full) {
                const int S0 = S;
                for (int s_ = std::max(2, S0 / 2); s_ <= std::min(128, 4 * S0); ++s_) {
                    bool f2;
                    const double c = cost(s_, f2);
                    if (f2 && c < best - 1e-9) { best = c; S = s_; }
                }
                // the model is flat (+-1 %) between 45 and 60 segments; the measured optimum of the real-symmetric kernels
                // on C3 is 50 segments (S = 20: profiles/r2_s11_c3_sweep.txt, r2_s15_c3_S.txt: 0.369 ms against 0.373 at 40 and 0.379 at 44 / 48
                // segments, three interleaved repetitions): take 50 when that still gives >= 2 waves of warps
                if (h->seg_real) {
                    const int s48 = (NT + 49) / 50;
                    if (s48 >= 2 && KGR * (((NT + s48 - 1) / s48 + SPW - 1) / SPW) >= 2 * resident) S = s48;
                }
            }
        }
        // enough (generator, segment) pairs to fill the GPU: one thread forms the propagators of its segment and their
        // product; GRAPE_B200_NO_FORMSEG=1 / GRAPE_B200_FORCE_FORMSEG=1 override the size rule (tests)
        h->seg_fuse = (long long)G * ((NT + S - 1) / S) >= 32768;
        // the scan schedule below runs one block of up to 128 segment threads per generator: worthwhile from 64 generators
        // on (a 512-trajectory shard of the C3 ensemble on 8 GPUs must not drop to the unfused kernels)
        if (h->seg_real && h->sym_v != 1 && G >= 64 && NT >= 64) h->seg_fuse = true;
        // ... and small ensembles down to the single README trajectory (C1): formation + scan, tau, gradient, finalize
        // are 4 short kernels in one CUDA graph instead of 9 (per-step formation, segment products, two chains, ...)
        if (h->seg_real && h->sym_v != 1 && K <= 2048 && NT >= 8) h->seg_fuse = true;
        if (getenv("GRAPE_B200_FORCE_FORMSEG") && atoi(getenv("GRAPE_B200_FORCE_FORMSEG")) != 0) h->seg_fuse = true;
        if (getenv("GRAPE_B200_NO_FORMSEG")) h->seg_fuse = false;
        // real-symmetric generators, fused formation: prefix products by a parallel scan inside the formation kernel
        // (one block of <= 128 segment threads per generator), no boundary chains (small_sym.cuh).  Without the chains
        // short segments cost nothing extra, and the measured optimum of formation + contraction is at the shortest
        // segments that keep the block full (profiles/r2_s1_c3_sweep.txt): NSEG ~ 128 for shards up to 2048
        // trajectories, ~ 64 above.  GRAPE_B200_SEG_SCAN=0 keeps the chains.
        // Measured (profiles/r2_s2_c3_sweep.txt): the scan's extra products (+10 % formation work) pay off when the
        // sequential chains are a visible share of the step, i.e. for shards of up to 2048 trajectories (0.133 ->
        // 0.103 ms at K = 512); at K = 4096 the chains are 7 % of the step and stay.  GRAPE_B200_SEG_SCAN=1 / 0 forces.
        h->seg_scan = h->seg_real && h->sym_v != 1 && h->seg_fuse && L <= 16 && K <= 2048;
        if (const char* e = getenv("GRAPE_B200_SEG_SCAN"))
            h->seg_scan = h->seg_real && h->sym_v != 1 && h->seg_fuse && L <= 16 && atoi(e) != 0;
        if (h->seg_scan) {
            const long long KGR = (K + a.BKL - 1) / a.BKL;
            // segments per generator: measured optimum 63 (K = 2048), 91 (K <= 1024); a handful of trajectories: as many
            // as one block holds (pure dependency latency, the shortest segments win)
            const int target = KGR >= 64 ? 64 : (KGR >= 8 ? 96 : 128);
            S = (NT + target - 1) / target;
            if (S < 2) S = 2;
        }
        if (h->seg_scan) {
            if (int rc = dev_alloc(h, &a.tau_part, (size_t)((K + 255) / 256) * 4)) return rc;
            if (int rc = dev_alloc(h, &a.tau_ticket, 1)) return rc;
            CUDA_TRY(h, cudaMemset(a.tau_ticket, 0, sizeof(int)));
            a.scan = 1;
        }
        if (const char* e = getenv("GRAPE_B200_SEG_S")) S = atoi(e);
        a.S = S < 2 ? 2 : (S > 128 ? 128 : S);
        a.NSEG = (NT + a.S - 1) / a.S;
        if (h->seg_scan && a.NSEG > 32 * SCAN_MAXW) {   // a forced S too short for one block per generator
            a.S = (NT + 32 * SCAN_MAXW - 1) / (32 * SCAN_MAXW);
            a.NSEG = (NT + a.S - 1) / a.S;
        }
        p.KB = (K + a.BKL - 1) / a.BKL;
        if (int rc = dev_alloc(h, &a.Pseg, (size_t)a.NSEG * NN * G)) return rc;
        if (int rc = dev_alloc(h, &a.chiE, (size_t)a.NSEG * N * K)) return rc;
    }
    if (int rc = dev_alloc(h, &p.partial, (size_t)p.KB * L * NT)) return rc;
    switch (N) {
        case 1: return small_set_attrs<1>(h);
        case 2: return small_set_attrs<2>(h);
        case 3: return small_set_attrs<3>(h);
        case 4: return small_set_attrs<4>(h);
    }
    h->err = "small path requires N <= 4";
    return GRAPE_B200_EINVAL;
}

template <int N>
void small_formU_t(H* h) {
    const long long tot = (long long)h->p.G * h->p.NT;
    small_form_U<N><<<(unsigned)((tot + 127) / 128), 128, 0, h->stream>>>(h->p);
    h->launches++;
}
template <int N>
void small_forward_t(H* h) {
    const int bd = small_bd(h);
    small_forward<N, SMALL_D><<<(h->p.K + bd - 1) / bd, bd, small_smem(h), h->stream>>>(h->p);
    h->launches++;
}
template <int N>
void small_backward_t(H* h, const cplx* chi_host) {
    const int bd = small_bd(h);
    small_backward<N, SMALL_D><<<(h->p.K + bd - 1) / bd, bd, small_smem(h), h->stream>>>(h->p, chi_host);
    h->launches++;
}
template <int N, int LCMAX>
void small_gradient_t(H* h) {
    const DevP& p = h->p;
    int BK = 1;
    while (BK < p.K && BK < 128) BK <<= 1;
    const int BN = 128 / BK;
    dim3 grid((p.K + BK - 1) / BK, (p.NT + BN - 1) / BN);
    int l0 = 0;
    while (l0 < p.L) {
        const int rem = p.L - l0;
        if (LCMAX >= 4 && rem >= 4) { small_gradient<N, (LCMAX >= 4 ? 4 : 1)><<<grid, 128, 0, h->stream>>>(p, l0, BK); l0 += 4; }
        else if (LCMAX >= 2 && rem >= 2) { small_gradient<N, (LCMAX >= 2 ? 2 : 1)><<<grid, 128, 0, h->stream>>>(p, l0, BK); l0 += 2; }
        else { small_gradient<N, 1><<<grid, 128, 0, h->stream>>>(p, l0, BK); l0 += 1; }
        h->launches++;
    }
}

// ---- time-segmented schedule (small_seg.cuh)
template <int N>
void seg_prod_t(H* h) {
    const long long tot = (long long)h->p.G * h->seg.NSEG;
    small_segprod<N><<<(unsigned)((tot + 127) / 128), 128, 0, h->stream>>>(h->p, h->seg);
    h->launches++;
}
template <int N>
void seg_formseg_t(H* h) {
    const long long tot = (long long)h->p.G * h->seg.NSEG;
    if (N <= 3 && h->seg_real) {
        constexpr int NS = N <= 3 ? N : 1;
        const unsigned blocks = (unsigned)((tot + SYM_BD - 1) / SYM_BD);
        const size_t sm = sym_stage_bytes(N, h->p.L);
        if (h->seg_scan) {   // one block per generator, thread = segment, prefix products by a warp scan
            const unsigned bd = 32u * (unsigned)((h->seg.NSEG + 31) / 32);
            if (h->p.L == 1) small_formscan_sym<NS, 1><<<h->p.G, bd, 0, h->stream>>>(h->p, h->seg);
            else if (h->p.L == 2) small_formscan_sym<NS, 2><<<h->p.G, bd, 0, h->stream>>>(h->p, h->seg);
            else small_formscan_sym<NS, 0><<<h->p.G, bd, 0, h->stream>>>(h->p, h->seg);
        } else if (h->sym_v == 1) {
            cudaMemsetAsync(h->seg.notfast, 0, sizeof(int), h->stream);
            small_formseg_sym<NS><<<(unsigned)((tot + 127) / 128), 128, 0, h->stream>>>(h->p, h->seg);
        } else if (h->p.L == 1) small_formseg_sym2<NS, 1><<<blocks, SYM_BD, sm, h->stream>>>(h->p, h->seg);
        else if (h->p.L == 2) small_formseg_sym2<NS, 2><<<blocks, SYM_BD, sm, h->stream>>>(h->p, h->seg);
        else small_formseg_sym2<NS, 0><<<blocks, SYM_BD, sm, h->stream>>>(h->p, h->seg);
    } else {
        small_formseg<N><<<(unsigned)((tot + 127) / 128), 128, 0, h->stream>>>(h->p, h->seg);
    }
    h->launches++;
}
template <int N>
void seg_scan_tau_t(H* h) {
    small_scan_tau<(N <= 3 ? N : 1)><<<(h->p.K + 255) / 256, 256, 0, h->stream>>>(h->p, h->seg);
    h->launches++;
}
template <int N>
void seg_scan_bounds_t(H* h, const cplx* chi_host, int fwd_only) {
    const long long tot = (long long)h->p.K * h->seg.NSEG;
    small_scan_bounds<(N <= 3 ? N : 1)><<<(unsigned)((tot + 127) / 128), 128, 0, h->stream>>>(h->p, h->seg, chi_host, fwd_only);
    h->launches++;
}
template <int N>
void seg_chain_fwd_t(H* h) {
    small_segchain_fwd<N><<<(h->p.K + 63) / 64, 64, 0, h->stream>>>(h->p, h->seg);
    h->launches++;
}
template <int N>
void seg_chain_dual_t(H* h) {
    small_segchain_dual<N><<<(h->p.K + CHAIN_BD - 1) / CHAIN_BD, 2 * CHAIN_BD, chain_ring_bytes(N), h->stream>>>(h->p, h->seg);
    h->launches++;
}
template <int N>
void seg_fwd_t(H* h) {
    const long long tot = (long long)h->p.K * h->seg.NSEG;
    small_segfwd<N><<<(unsigned)((tot + 127) / 128), 128, 0, h->stream>>>(h->p, h->seg);
    h->launches++;
}
template <int N>
void seg_chain_bwd_t(H* h, const cplx* chi_host) {
    small_segchain_bwd<N><<<(h->p.K + 63) / 64, 64, 0, h->stream>>>(h->p, h->seg, chi_host);
    h->launches++;
}
// the fused formation + segment-product kernel runs (enough (generator, segment) pairs to fill the GPU)
bool seg_fused(const H* h) { return h->seg_on && h->seg_fuse; }
// the real-symmetric gradient kernel may serve this call: its eligibility flag was written by small_formseg_sym
bool seg_real_active(const H* h) {
    return h->seg_real && seg_fused(h) && h->p.grad_method == 0 && !h->p.taugrads;
}
// gradient calls served by the staged real-symmetric kernel on the chain schedule: both boundary chains in one pass
bool chain_dual_ok(const H* h) {
    return h->seg_on && !h->seg_scan && h->p.N <= 3 && h->sym_v != 1 && seg_real_active(h) &&
           h->p.functional != GRAPE_B200_JT_HOST &&
           !(getenv("GRAPE_B200_CHAIN_DUAL") && atoi(getenv("GRAPE_B200_CHAIN_DUAL")) == 0);
}
template <int N, int LCMAX, bool HERM>
void seg_grad_launch(H* h, const int* run_if) {
    const DevP& p = h->p;
    const SegArgs& a = h->seg;
    const int SPW = 32 / a.BKL;
    const long long warps = (long long)((p.K + a.BKL - 1) / a.BKL) * ((a.NSEG + SPW - 1) / SPW);
    const unsigned blocks = (unsigned)((warps + 3) / 4);
    int l0 = 0;
    while (l0 < p.L) {
        const int rem = p.L - l0;
        if (LCMAX >= 4 && rem >= 4) { small_seggrad<N, (LCMAX >= 4 ? 4 : 1), HERM><<<blocks, 128, 0, h->stream>>>(p, a, l0, run_if); l0 += 4; }
        else if (LCMAX >= 2 && rem >= 2) { small_seggrad<N, (LCMAX >= 2 ? 2 : 1), HERM><<<blocks, 128, 0, h->stream>>>(p, a, l0, run_if); l0 += 2; }
        else { small_seggrad<N, 1, HERM><<<blocks, 128, 0, h->stream>>>(p, a, l0, run_if); l0 += 1; }
        h->launches++;
    }
}
template <int N, int LCMAX>
void seg_grad_t(H* h) {
    const int* run_if = nullptr;
    if (N <= 3 && seg_real_active(h)) {
        // real-symmetric generators: all controls in one launch
        const SegArgs& a = h->seg;
        const int SPW = 32 / a.BKL;
        const long long warps = (long long)((h->p.K + a.BKL - 1) / a.BKL) * ((a.NSEG + SPW - 1) / SPW);
        constexpr int NS = N <= 3 ? N : 1;
        const unsigned blocks = (unsigned)((warps + 3) / 4);
        const size_t sm = sym_stage_bytes(N, h->p.L);
        if (h->sym_v == 1) {
            // round-1 kernels (operators re-loaded from global memory every step): if a step of this call is not eligible
            // (flag written by small_formseg_sym) the kernel returns at once and the general Hermitian kernel below runs
            if (h->sym_occ == 2) small_seggrad_sym<NS, 2><<<blocks, 128, 0, h->stream>>>(h->p, a);
            else small_seggrad_sym<NS, 3><<<blocks, 128, 0, h->stream>>>(h->p, a);
            h->launches++;
            run_if = a.notfast;
        } else {
            // staged kernels: every step is served (sub-stepping inside), no second launch
            // (two-warp blocks for under-filled launches were measured: no difference, profiles/r2_s13_c3_sweep.txt)
            // one or two controls: 128 registers (16 warps / SM), orders 7 and 8 out of line (profiles/r2_s19_c3_variants.txt:
            // 0.311 ms against 0.322 ms with 168 registers and all orders inline; GRAPE_B200_SYM_OCC=10 selects the latter)
            if (h->p.L == 1) small_seggrad_sym2<NS, 1, 4, SYM_BD, false, 6><<<blocks, SYM_BD, sm, h->stream>>>(h->p, a);
            else if (h->p.L == 2 && h->sym_occ == 10) small_seggrad_sym2<NS, 2, 3, SYM_BD, false, 8><<<blocks, SYM_BD, sm, h->stream>>>(h->p, a);
            else if (h->p.L == 2) small_seggrad_sym2<NS, 2, 4, SYM_BD, false, 6><<<blocks, SYM_BD, sm, h->stream>>>(h->p, a);
            else small_seggrad_sym2<NS, 0><<<blocks, SYM_BD, sm, h->stream>>>(h->p, a);
            h->launches++;
            return;
        }
    }
    if (h->seg_scan && !h->chib_done) {   // general kernel after a fast-path backward call (tau_grads dump)
        seg_scan_bounds_t<N>(h, h->seg.chi_host, 0);
        h->bounds_done = h->chib_done = true;
    } else if (h->seg_on && !h->seg_scan && !h->chib_done) {   // ... after a one-pass chain call: the normalised chi chain
        seg_chain_bwd_t<N>(h, nullptr);
        h->seg.scan = 0;
        h->chib_done = true;
    }
    if (h->seg_herm) seg_grad_launch<N, LCMAX, true>(h, run_if);
    else seg_grad_launch<N, LCMAX, false>(h, run_if);
}

#define SMALL_DISPATCH(N_, CALL1, CALL2, CALL3, CALL4) \
    switch (N_) { case 1: CALL1; break; case 2: CALL2; break; case 3: CALL3; break; default: CALL4; break; }

// ------------------------------------------------------------------ phase drivers
void run_formU(H* h) {
    switch (h->path) {
        case GRAPE_B200_PATH_SMALL:
            if (seg_fused(h)) {
                // enough (generator, segment) pairs to fill the GPU: fused formation + segment product
                SMALL_DISPATCH(h->p.N, seg_formseg_t<1>(h), seg_formseg_t<2>(h), seg_formseg_t<3>(h), seg_formseg_t<4>(h));
                h->U_valid = h->seg.store_U != 0;
                break;
            }
            SMALL_DISPATCH(h->p.N, small_formU_t<1>(h), small_formU_t<2>(h), small_formU_t<3>(h), small_formU_t<4>(h));
            h->U_valid = true;
            if (h->seg_on) { SMALL_DISPATCH(h->p.N, seg_prod_t<1>(h), seg_prod_t<2>(h), seg_prod_t<3>(h), seg_prod_t<4>(h)); }
            break;
        case GRAPE_B200_PATH_WARP:
            warp_run_formU(h->warp, h->p, h->stream, h->launches);
            if (h->wseg_on) warp_seg_run_prod(h->wseg, h->warp, h->p, h->stream, h->launches);
            break;
        case GRAPE_B200_PATH_DENSE: break;   // dense path never forms U
    }
}
void run_fill_interior(H* h) {
    if (h->path == GRAPE_B200_PATH_SMALL && h->seg_on && !h->interior_done) {
        if (h->seg_scan && !h->bounds_done) {   // functional-only call: the boundary states were never needed so far
            SMALL_DISPATCH(h->p.N, seg_scan_bounds_t<1>(h, nullptr, 1), seg_scan_bounds_t<2>(h, nullptr, 1),
                           seg_scan_bounds_t<3>(h, nullptr, 1), seg_scan_bounds_t<3>(h, nullptr, 1));
            h->bounds_done = true;
        }
        if (!h->U_valid) {   // Hermitian schedule: the propagators were only accumulated into the segment products
            SMALL_DISPATCH(h->p.N, small_formU_t<1>(h), small_formU_t<2>(h), small_formU_t<3>(h), small_formU_t<4>(h));
            h->U_valid = true;
        }
        SMALL_DISPATCH(h->p.N, seg_fwd_t<1>(h), seg_fwd_t<2>(h), seg_fwd_t<3>(h), seg_fwd_t<4>(h));
        h->interior_done = true;
    }
    if (h->path == GRAPE_B200_PATH_WARP && h->wseg_on && !h->interior_done) {
        warp_seg_run_fill(h->wseg, h->warp, h->p, h->stream, h->launches);
        h->interior_done = true;
    }
}
bool xchg_live(const H* h) { return h->xchg_on && h->xchg_mode; }
// start of an evaluation: remember whether it is a functional-only call (its sums are always exchanged)
void run_xchg_begin(H* h, bool grad) { h->xchg_fonly = !grad; }
void run_forward(H* h, bool need_storage = true, bool with_backward = false) {
    switch (h->path) {
        case GRAPE_B200_PATH_SMALL:
            if (h->seg_on && h->seg_scan) {   // tau, final states and the sums straight from the prefix products
                SMALL_DISPATCH(h->p.N, seg_scan_tau_t<1>(h), seg_scan_tau_t<2>(h), seg_scan_tau_t<3>(h), seg_scan_tau_t<3>(h));
                h->interior_done = false;
                h->bounds_done = false;
                h->chib_done = false;
                break;
            }
            if (h->seg_on) {
                // gradient call served by the staged real-symmetric kernel: both boundary chains in one pass
                h->chain_dual = with_backward && chain_dual_ok(h);
                if (h->chain_dual) {
                    SMALL_DISPATCH(h->p.N, seg_chain_dual_t<1>(h), seg_chain_dual_t<2>(h), seg_chain_dual_t<3>(h), seg_chain_dual_t<3>(h));
                } else {
                    SMALL_DISPATCH(h->p.N, seg_chain_fwd_t<1>(h), seg_chain_fwd_t<2>(h), seg_chain_fwd_t<3>(h), seg_chain_fwd_t<4>(h));
                }
                h->interior_done = false;
                if (need_storage && !h->seg_herm) run_fill_interior(h);
                break;
            }
            SMALL_DISPATCH(h->p.N, small_forward_t<1>(h), small_forward_t<2>(h), small_forward_t<3>(h), small_forward_t<4>(h));
            break;
        case GRAPE_B200_PATH_WARP:
            if (h->wseg_on) {
                warp_seg_run_forward(h->wseg, h->warp, h->p, false, h->stream, h->launches);
                h->interior_done = false;
                if (need_storage) run_fill_interior(h);
            } else warp_run_forward(h->warp, h->p, h->stream, h->launches);
            break;
        case GRAPE_B200_PATH_DENSE:
            kry_run_plan(h->dense, h->p, h->stream, h->launches);
            if (h->dense2.on) dense2_run_forward(h->dense2, h->dense, h->p, h->stream, h->launches);
            else dense_run_forward(h->dense, h->p, h->stream, h->launches, with_backward);
            break;
    }
    // J_T_re / J_T_ss on one device: chi_k only needs tau_k, the sums are formed by the finalize kernel (one launch less)
    h->defer_tau = with_backward && h->p.functional != GRAPE_B200_JT_SM && h->p.functional != GRAPE_B200_JT_HOST &&
                   h->p.gb_kind == 0;
    if (!(h->path == GRAPE_B200_PATH_SMALL && h->seg_on && h->seg_scan) && !h->defer_tau) {   // (the scan schedule's tau kernel has done it)
        reduce_tau<<<1, h->p.K >= 2048 ? 1024 : 256, 0, h->stream>>>(h->p);   // single block, fixed order; wider for large ensembles
        h->launches++;
    }
    // sharded over peers: chi of J_T_sm needs the global sum of tau before the backward sweep (optimize.jl:845-855);
    // a functional-only call needs the global sums for J itself. Otherwise the sums travel with the gradient.
    if (xchg_live(h) && (h->xchg_fonly || h->p.functional == GRAPE_B200_JT_SM)) {
        xchg_sums<<<1, 32, 0, h->stream>>>(h->p, h->xd);
        h->launches++;
    }
}
void run_backward(H* h, const cplx* chi_host) {
    switch (h->path) {
        case GRAPE_B200_PATH_SMALL:
            if (h->seg_on && h->seg_scan) {
                // the real-symmetric gradient kernel derives its boundary states from the prefix products itself; the
                // general kernels (:taylor, tau_grads dump) read the arrays small_scan_bounds fills
                h->seg.chi_host = chi_host;
                h->seg.scan = 1;
                h->chib_done = false;
                if (!seg_real_active(h)) {
                    SMALL_DISPATCH(h->p.N, seg_scan_bounds_t<1>(h, chi_host, 0), seg_scan_bounds_t<2>(h, chi_host, 0),
                                   seg_scan_bounds_t<3>(h, chi_host, 0), seg_scan_bounds_t<3>(h, chi_host, 0));
                    h->bounds_done = h->chib_done = true;
                }
                break;
            }
            if (h->seg_on && h->chain_dual && !chi_host && seg_real_active(h)) {
                h->seg.scan = 2;     // chiE already holds the propagated targets; the gradient kernel applies c_k / rho_k
                h->chib_done = false;
                break;
            }
            if (h->seg_on) {
                h->seg.scan = 0;
                h->chib_done = true;
                if (!h->seg_herm) run_fill_interior(h);
                SMALL_DISPATCH(h->p.N, seg_chain_bwd_t<1>(h, chi_host), seg_chain_bwd_t<2>(h, chi_host),
                               seg_chain_bwd_t<3>(h, chi_host), seg_chain_bwd_t<4>(h, chi_host));
                break;
            }
            SMALL_DISPATCH(h->p.N, small_backward_t<1>(h, chi_host), small_backward_t<2>(h, chi_host),
                           small_backward_t<3>(h, chi_host), small_backward_t<4>(h, chi_host));
            break;
        case GRAPE_B200_PATH_WARP:
            if (h->wseg_on) {
                run_fill_interior(h);
                warp_seg_run_backward(h->wseg, h->warp, h->p, chi_host, h->stream, h->launches);
            } else warp_run_backward(h->warp, h->p, chi_host, h->stream, h->launches);
            break;
        case GRAPE_B200_PATH_DENSE:
            if (h->dense2.on) dense2_run_backward(h->dense2, h->dense, h->p, chi_host, h->stream, h->launches);
            else dense_run_backward(h->dense, h->p, chi_host, h->stream, h->launches);
            break;
    }
}
void run_gradient(H* h) {
    switch (h->path) {
        case GRAPE_B200_PATH_SMALL:
            if (h->seg_on) {
                SMALL_DISPATCH(h->p.N, (seg_grad_t<1, 4>(h)), (seg_grad_t<2, 4>(h)), (seg_grad_t<3, 2>(h)), (seg_grad_t<4, 1>(h)));
                break;
            }
            SMALL_DISPATCH(h->p.N, (small_gradient_t<1, 4>(h)), (small_gradient_t<2, 4>(h)),
                           (small_gradient_t<3, 2>(h)), (small_gradient_t<4, 1>(h)));
            break;
        case GRAPE_B200_PATH_WARP: warp_run_gradient(h->warp, h->p, h->stream, h->launches); break;
        case GRAPE_B200_PATH_DENSE:   // block recursion: fused into the backward sweep; Krylov form: contraction over all steps
            kry_run_gradient(h->dense, h->p, h->stream, h->launches, h->dense.dual_launched);
            break;
    }
}
void run_finalize(H* h, bool grad) {
    if (grad) {
        const int blocks = (h->LNT + 31) / 32;
        if (xchg_live(h)) {  // k-reduction fused with the all-reduce over the peers' shards (NVLink stores, xchg.cuh)
            finalize_grad_xchg<<<blocks < XCHG_MAXB ? blocks : XCHG_MAXB, 256, 0, h->stream>>>(
                h->p, h->xd, h->p.functional != GRAPE_B200_JT_SM ? 1 : 0, h->defer_tau ? 1 : 0);
            h->launches++;
            h->defer_tau = false;
            return;
        } else {             // k-reduction, J_a gradient and (last block) J_parts in one launch
            finalize_grad<<<blocks < 592 ? blocks : 592, 256, 0, h->stream>>>(h->p, 1, h->defer_tau ? 1 : 0);
            h->launches++;
            h->defer_tau = false;
            return;
        }
    }
    finalize_J<<<1, 256, 0, h->stream>>>(h->p, 0);
    h->launches++;
}
void rec(H* h, int i) {
    if (h->profiling) cudaEventRecord(h->ev[i], h->stream);
}
void collect_timings(H* h, int64_t launches_before) {
    if (!h->profiling) return;
    float ms;
    for (int i = 0; i < 6; ++i) {
        ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
        h->timings[i] = ms;
    }
    ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]);
    h->timings[6] = ms;
    h->timings[7] = (double)(h->launches - launches_before);
}

int check_flags(H* h) {
    const DevFlags* f = reinterpret_cast<const DevFlags*>(h->h_out + h->off_flags);
    char b[256];
    if (f->xchg_timeout == 2) {
        h->err = "internal error: a split-phase barrier of the concurrent dense chains timed out";
        return GRAPE_B200_ECUDA;
    }
    if (f->xchg_timeout) {
        h->err = "peer exchange timed out: a shard of this trajectory-sharded problem did not reach the exchange "
                 "(every rank must make the same sequence of evaluation calls)";
        return GRAPE_B200_ENCCL;
    }
    if (f->chi_bad_k) {
        // message of reference src/optimize.jl:1021-1025
        snprintf(b, sizeof b, "The χ state with index %d has norm %g < %g (chi_min_norm)",
                 f->chi_bad_k, f->chi_bad_rho, h->p.chi_min_norm);
        h->err = b;
        return GRAPE_B200_ECHINORM;
    }
    if (f->taylor_fail) {
        // message of reference src/optimize.jl:644-648
        snprintf(b, sizeof b, "taylor_grad_step! did not converge within %d iterations. Residual term r=%g.",
                 h->p.taylor_max_order, f->taylor_r);
        h->err = b;
        return GRAPE_B200_ETAYLOR;
    }
    return 0;
}

int upload_pulses(H* h, const double* pulsevals) {
    memcpy(h->h_in, pulsevals, sizeof(double) * h->LNT);
    h->p.eps = h->d_eps_own;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_eps_own, h->h_in, sizeof(double) * h->LNT,
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->p.flags, 0, sizeof(DevFlags), h->stream));
    return 0;
}
int download_enqueue(H* h) {
    CUDA_TRY(h, cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * h->out_doubles,
                                cudaMemcpyDeviceToHost, h->stream));
    rec(h, 6);
    return 0;
}
int download_wait(H* h) {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}
int download_all(H* h) {
    if (int rc = download_enqueue(h)) return rc;
    return download_wait(h);
}

// Whole-call CUDA graph (small / sub-warp paths, profiling off). Returns 1 if the call was enqueued as a graph
// launch, 0 if the caller must launch directly, < 0 on error (code negated). No synchronisation.
int eval_via_graph(H* h, const double* pulsevals, bool grad) {
    if (!h->graphs_ok || h->profiling || h->path == GRAPE_B200_PATH_DENSE) return 0;
    cudaGraphExec_t& ge = grad ? h->graph_fg : h->graph_f;
    int64_t& gl = grad ? h->graph_fg_launches : h->graph_f_launches;
    memcpy(h->h_in, pulsevals, sizeof(double) * h->LNT);
    h->p.eps = h->d_eps_own;
    if (!ge) {
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError(); h->graphs_ok = false; return 0;
        }
        const int64_t l0 = h->launches;
        cudaMemcpyAsync(h->d_eps_own, h->h_in, sizeof(double) * h->LNT, cudaMemcpyHostToDevice, h->stream);
        cudaMemsetAsync(h->p.flags, 0, sizeof(DevFlags), h->stream);
        run_xchg_begin(h, grad);
        run_formU(h);
        run_forward(h, grad, grad);
        if (grad) { run_backward(h, nullptr); run_gradient(h); }
        run_finalize(h, grad);
        cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * h->out_doubles, cudaMemcpyDeviceToHost, h->stream);
        gl = h->launches - l0;
        h->launches = l0;
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        if (e == cudaSuccess) e = cudaGraphInstantiate(&ge, g, 0);
        if (g) cudaGraphDestroy(g);
        if (e != cudaSuccess) { cudaGetLastError(); ge = nullptr; h->graphs_ok = false; return 0; }
    }
    h->xchg_fonly = !grad;
    if (cudaGraphLaunch(ge, h->stream) != cudaSuccess) { cudaGetLastError(); h->graphs_ok = false; return 0; }
    h->launches += gl;
    // the captured sequence of eval_f leaves the interior of fw_storage unfilled
    if (h->seg_on || h->wseg_on) h->interior_done = grad && !h->seg_herm;
    if (h->seg_scan) h->bounds_done = h->chib_done = grad && !seg_real_active(h);
    else if (h->seg_on) { h->chain_dual = grad && chain_dual_ok(h); h->chib_done = grad && !h->chain_dual; }
    if (h->path == GRAPE_B200_PATH_SMALL)   // same bookkeeping as run_formU (not executed on a graph replay)
        h->U_valid = !(seg_fused(h) && !h->seg.store_U);
    return 1;
}
void drop_graphs(H* h) {
    if (h->graph_fg) { cudaGraphExecDestroy(h->graph_fg); h->graph_fg = nullptr; }
    if (h->graph_f) { cudaGraphExecDestroy(h->graph_f); h->graph_f = nullptr; }
}

// One complete evaluation (evaluate_functional / evaluate_gradient!, reference src/optimize.jl:696-768, 824-1014),
// enqueued on the handle's stream without any host synchronisation: H2D of the pulses, the kernel sequence
// (with the peer exchanges if shards are attached), D2H of the output block.  eval_wait() synchronises.
int eval_launch(H* h, const double* pulsevals, bool grad) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    h->xchg_mode = true;
    h->launch_l0 = h->launches;
    const int via = eval_via_graph(h, pulsevals, grad);
    if (via < 0) return -via;
    h->launched_via_graph = via == 1;
    if (!via) {
        rec(h, 0);
        if (int rc = upload_pulses(h, pulsevals)) return rc;
        run_xchg_begin(h, grad);
        run_formU(h); rec(h, 1);
        run_forward(h, grad, grad); rec(h, 2);   // functional only: the interior of fw_storage is filled lazily (get_stored_states)
        rec(h, 3);
        if (grad) { run_backward(h, nullptr); rec(h, 4); run_gradient(h); }
        else rec(h, 4);
        run_finalize(h, grad); rec(h, 5);
        if (int rc = download_enqueue(h)) return rc;
    }
    return 0;
}
int eval_wait(H* h, bool grad) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (cudaStreamSynchronize(h->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        h->err = "CUDA error while executing the evaluation";
        return GRAPE_B200_ECUDA;
    }
    if (!h->launched_via_graph) collect_timings(h, h->launch_l0);
    h->forward_done = true; h->backward_done = grad; h->taugrads_valid = false;
    return 0;
}

}  // namespace

struct grape_b200_handle : grape_b200_handle_impl {};

extern "C" {

int grape_b200_abi_version(void) { return GRAPE_B200_ABI_VERSION; }

const char* grape_b200_last_error(const grape_b200_handle* h) {
    return h ? h->err.c_str() : g_create_error.c_str();
}

void grape_b200_destroy(grape_b200_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->graph_fg) cudaGraphExecDestroy(h->graph_fg);
    if (h->graph_f) cudaGraphExecDestroy(h->graph_f);
    for (void* q : h->xchg_ipc_opened) cudaIpcCloseMemHandle(q);
    if (h->xchg_buf) cudaFree(h->xchg_buf);
    dense_destroy(h->dense);
    for (void* q : h->dev_allocs) cudaFree(q);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_in) cudaFreeHost(h->h_in);
    for (int i = 0; i < 8; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int grape_b200_create(const grape_b200_problem* d, grape_b200_handle** out) {
    if (!out) return GRAPE_B200_EINVAL;
    *out = nullptr;
    auto fail = [&](int code, const std::string& msg) { g_create_error = msg; return code; };
    if (!d) return fail(GRAPE_B200_EINVAL, "null problem descriptor");
    if (d->abi_version != GRAPE_B200_ABI_VERSION) return fail(GRAPE_B200_EINVAL, "ABI version mismatch");
    if (d->L <= 0) return fail(GRAPE_B200_ENOCONTROLS, "no controls in trajectories: cannot optimize");
    if (d->K <= 0 || d->N <= 0 || d->NT <= 0 || d->G <= 0 || d->G > d->K)
        return fail(GRAPE_B200_EINVAL, "invalid dimensions (need K,N,NT > 0 and 1 <= G <= K)");
    if (!d->tlist || !d->H0 || !d->Hc || !d->psi0 || !d->tgt)
        return fail(GRAPE_B200_EINVAL, "tlist, H0, Hc, psi0 and tgt are required");
    if (!d->gen_of_traj && d->G != 1 && d->G != d->K)
        return fail(GRAPE_B200_EINVAL, "gen_of_traj is required unless G == 1 or G == K");
    if (d->gb_kind != GRAPE_B200_GB_NONE && (!d->gb_D || (d->gb_nD != 1 && d->gb_nD != d->K)))
        return fail(GRAPE_B200_EINVAL, "g_b quadratic form needs gb_D with gb_nD == 1 or K");
    if (d->functional < 0 || d->functional > 3 || d->gradient_method < 0 || d->gradient_method > 1)
        return fail(GRAPE_B200_EINVAL, "invalid functional / gradient_method");
    for (int n = 0; n < d->NT; ++n)
        if (!(d->tlist[n + 1] > d->tlist[n])) return fail(GRAPE_B200_EINVAL, "tlist must be strictly increasing");

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(GRAPE_B200_ECUDA, std::string("no CUDA device available (") +
                                          cudaGetErrorString(ce) + "); this engine has no CPU fallback");
    if (d->device < 0 || d->device >= ndev) return fail(GRAPE_B200_EINVAL, "invalid device ordinal");

    grape_b200_handle* h = new grape_b200_handle();
    memset(&h->p, 0, sizeof(DevP));
    h->d_out = nullptr; h->h_out = nullptr; h->h_in = nullptr; h->d_chi_host = nullptr; h->d_damp = nullptr;
    h->d_tmp = nullptr; h->tmp_elems = 0; h->stream = nullptr;
    for (int i = 0; i < 8; ++i) { h->ev[i] = nullptr; h->timings[i] = 0.0; }
    h->profiling = false; h->forward_done = false; h->backward_done = false; h->taugrads_valid = false; h->launches = 0;
    h->seg_on = false; h->interior_done = false; memset(&h->seg, 0, sizeof h->seg);
    h->seg_herm = false; h->U_valid = false; h->seg_real = false; h->seg_fuse = false; h->sym_occ = 3; h->sym_v = 2; h->seg_scan = false; h->bounds_done = false; h->chib_done = false; h->chain_dual = false; h->defer_tau = false; h->d_taugrads = nullptr; h->taugrads_valid = false;
    h->wseg_on = false; memset(&h->wseg, 0, sizeof h->wseg);
    memset(&h->xd, 0, sizeof h->xd); h->xchg_on = false; h->xchg_mode = false; h->xchg_fonly = false;
    h->xchg_buf = nullptr; h->xchg_bytes = 0; h->launched_via_graph = false; h->launch_l0 = 0;
    h->graph_fg = nullptr; h->graph_f = nullptr; h->graph_fg_launches = h->graph_f_launches = 0;
    h->graphs_ok = !(getenv("GRAPE_B200_NO_GRAPH") && atoi(getenv("GRAPE_B200_NO_GRAPH")) != 0);
    h->device = d->device;
    auto bail = [&](int rc) {
        g_create_error = h->err;
        grape_b200_destroy(h);
        return rc;
    };
#define TRYC(call) do { int _rc = (call); if (_rc) return bail(_rc); } while (0)
#define CUDA_TRYC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { h->err = std::string("CUDA error: ") + cudaGetErrorString(_e); return bail(GRAPE_B200_ECUDA); } } while (0)
    CUDA_TRYC(cudaSetDevice(d->device));
    CUDA_TRYC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) CUDA_TRYC(cudaEventCreate(&h->ev[i]));

    DevP& p = h->p;
    p.K = d->K; p.N = d->N; p.L = d->L; p.NT = d->NT; p.G = d->G;
    p.Kglobal = d->K_global > 0 ? d->K_global : d->K;
    p.functional = d->functional; p.grad_method = d->gradient_method;
    p.ja_kind = d->ja_kind; p.gb_kind = d->gb_kind; p.gb_nD = d->gb_kind ? d->gb_nD : 1;
    p.taylor_max_order = d->taylor_max_order > 0 ? d->taylor_max_order : 100;
    p.taylor_check = d->taylor_check_convergence;
    p.taylor_tol = d->taylor_tolerance > 0 ? d->taylor_tolerance : 1e-16;
    p.lambda_a = d->lambda_a; p.lambda_b = d->lambda_b;
    p.chi_min_norm = d->chi_min_norm > 0 ? d->chi_min_norm : 1e-100;
    h->LNT = p.L * p.NT;
    h->path = choose_path(p.N, d->path);
    if (h->path == GRAPE_B200_PATH_SMALL && p.N > 4) { h->err = "PATH_SMALL requires N <= 4"; return bail(GRAPE_B200_EINVAL); }
    if (h->path == GRAPE_B200_PATH_WARP && p.N > WARP_MAX_N) { h->err = "PATH_WARP requires N <= 32"; return bail(GRAPE_B200_EINVAL); }

    const int K = p.K, LNT = h->LNT;
    {
        double* q;
        TRYC(dev_upload(h, &q, d->tlist, (size_t)p.NT + 1)); p.tlist = q;
        TRYC(dev_alloc(h, &h->d_eps_own, (size_t)LNT)); p.eps = h->d_eps_own;
        if (d->shape) { TRYC(dev_upload(h, &q, d->shape, (size_t)LNT)); p.shape = q; p.dshape = q; }
        if (d->weights) { TRYC(dev_upload(h, &q, d->weights, (size_t)K)); p.w = q; }
        std::vector<int> gen(K);
        for (int k = 0; k < K; ++k) {
            gen[k] = d->gen_of_traj ? d->gen_of_traj[k] : (d->G == 1 ? 0 : k);
            if (gen[k] < 0 || gen[k] >= d->G) { h->err = "gen_of_traj entry out of range"; return bail(GRAPE_B200_EINVAL); }
        }
        int* gi;
        TRYC(dev_upload(h, &gi, gen.data(), (size_t)K)); p.gen = gi;
    }
    // output block
    h->off_J = (size_t)3 * LNT;
    h->off_sums = h->off_J + 3;
    h->off_tau = (h->off_sums + 4 + 1) & ~(size_t)1;
    h->off_flags = h->off_tau + 2 * (size_t)K;
    h->out_doubles = h->off_flags + (sizeof(DevFlags) + 7) / 8;
    TRYC(dev_alloc(h, &h->d_out, h->out_doubles));
    CUDA_TRYC(cudaMemset(h->d_out, 0, sizeof(double) * h->out_doubles));
    p.grad = h->d_out;
    p.Jparts = h->d_out + h->off_J;
    p.sums = h->d_out + h->off_sums;
    p.tau = reinterpret_cast<cplx*>(h->d_out + h->off_tau);
    p.flags = reinterpret_cast<DevFlags*>(h->d_out + h->off_flags);
    CUDA_TRYC(cudaMallocHost(&h->h_out, sizeof(double) * h->out_doubles));
    CUDA_TRYC(cudaMallocHost(&h->h_in, sizeof(double) * (LNT + 8)));
    TRYC(dev_alloc(h, &p.rho, (size_t)K));
    TRYC(dev_alloc(h, &p.jb, (size_t)K));
    CUDA_TRYC(cudaMemset(p.jb, 0, sizeof(double) * K));
    TRYC(dev_alloc(h, &p.chiT, (size_t)K * p.N));
    TRYC(dev_alloc(h, &h->d_chi_host, (size_t)K * p.N));

    int rc = 0;
    switch (h->path) {
        case GRAPE_B200_PATH_SMALL: rc = small_setup(h, d); break;
        case GRAPE_B200_PATH_WARP: {
            std::string e;
            rc = warp_setup(h->warp, p, d, h->dev_allocs, e);
            // time-segmented schedule unless a state running cost couples chi to Psi at every step
            h->wseg_on = !rc && p.gb_kind == 0 && d->path != GRAPE_B200_PATH_WARP_CHAIN;
            if (h->wseg_on) rc = warp_seg_setup(h->wseg, h->warp, p, d, h->dev_allocs, e);
            if (rc) h->err = e;
            break;
        }
        case GRAPE_B200_PATH_DENSE: {
            std::string e;
            rc = dense_setup(h->dense, p, d, h->dev_allocs, e);
            if (!rc) rc = dense2_setup(h->dense2, h->dense, p, h->dev_allocs, e);
            if (!rc && !h->dense2.on && !h->dense.strip_ok) { e = h->dense.strip_err; rc = GRAPE_B200_EINVAL; }
            if (!rc) rc = kry_setup(h->dense, h->dense2, p, h->dev_allocs, e);
            if (!rc) rc = dense_dual_setup(h->dense, p, h->dense2.on, h->dev_allocs, e);
            if (!rc) rc = dense_concurrent_setup(h->dense, p, d->tgt, h->dense2.on, h->dev_allocs, e);
            if (!rc) rc = dense2_multi_setup(h->dense2, h->dense, e);
            if (rc) h->err = e;
            break;
        }
        default: h->err = "unknown path"; rc = GRAPE_B200_EINVAL;
    }
    if (rc) return bail(rc);
    CUDA_TRYC(cudaDeviceSynchronize());
    *out = h;
    return GRAPE_B200_OK;
#undef TRYC
#undef CUDA_TRYC
}

static void copy_out_common(grape_b200_handle* h, double* J_parts, double* tau) {
    if (J_parts) memcpy(J_parts, h->h_out + h->off_J, 3 * sizeof(double));
    if (tau) memcpy(tau, h->h_out + h->off_tau, 2 * sizeof(double) * h->p.K);
}

int grape_b200_eval_f(grape_b200_handle* h, const double* pulsevals, double* J_parts, double* tau) {
    if (!h || !pulsevals) return GRAPE_B200_EINVAL;
    if (int rc = eval_launch(h, pulsevals, false)) return rc;
    if (int rc = eval_wait(h, false)) return rc;
    if (h->p.functional == GRAPE_B200_JT_HOST) h->h_out[h->off_J] = 0.0 / 0.0;
    copy_out_common(h, J_parts, tau);
    return check_flags(h);
}

static int eval_fg_copy_out(grape_b200_handle* h, double* G, double* J_parts, double* tau, double* grad_J_Tb,
                            double* grad_J_a) {
    const int LNT = h->LNT;
    if (G) memcpy(G, h->h_out, sizeof(double) * LNT);
    if (grad_J_Tb) memcpy(grad_J_Tb, h->h_out + LNT, sizeof(double) * LNT);
    if (grad_J_a) memcpy(grad_J_a, h->h_out + 2 * LNT, sizeof(double) * LNT);
    copy_out_common(h, J_parts, tau);
    return check_flags(h);
}

int grape_b200_eval_fg(grape_b200_handle* h, const double* pulsevals, double* G, double* J_parts,
                       double* tau, double* grad_J_Tb, double* grad_J_a) {
    if (!h || !pulsevals || !G) return GRAPE_B200_EINVAL;
    if (h->p.functional == GRAPE_B200_JT_HOST) {
        h->err = "eval_fg needs a built-in functional; use forward + backward_chi for JT_HOST";
        return GRAPE_B200_EINVAL;
    }
    if (int rc = eval_launch(h, pulsevals, true)) return rc;
    if (int rc = eval_wait(h, true)) return rc;
    return eval_fg_copy_out(h, G, J_parts, tau, grad_J_Tb, grad_J_a);
}

// ---- amplitude mode: non-linear controls / per-term amplitudes (reference src/workspace.jl:283-285 get_control_derivs,
// src/optimize.jl:946-951 incl. the isnothing(mu) -> 0 branch).  The descriptor's L "controls" are amplitude SLOTS:
// H_n = H0 + sum_i ampl[i][n] Hc_i and mu_{i,n} = dampl[i][n] Hc_i, both evaluated by the host from its closures
// (L*NT scalar evaluations per call).  G_slots[i][n] is the gradient with respect to the control value behind slot i
// THROUGH that slot; the host adds the slots of one control.  Direct launches (no graph: the kernel parameters differ).
static int eval_amplitudes(grape_b200_handle* h, const double* ampl, const double* dampl, bool grad) {
    if (h->p.ja_kind != GRAPE_B200_JA_NONE) {
        h->err = "amplitude mode: J_a acts on the control values, which the library does not see -- evaluate it on the host "
                 "(ja_kind must be GRAPE_B200_JA_NONE)";
        return GRAPE_B200_EINVAL;
    }
    if (h->p.shape) {
        h->err = "amplitude mode: the handle must be created without a static shape (fold it into ampl / dampl)";
        return GRAPE_B200_EINVAL;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int LNT = h->LNT;
    if (grad && !h->d_damp)
        if (int rc = dev_alloc(h, &h->d_damp, (size_t)LNT)) return rc;
    h->xchg_mode = true;
    h->launch_l0 = h->launches;
    h->launched_via_graph = false;
    const double* shape0 = h->p.shape;
    const double* dshape0 = h->p.dshape;
    rec(h, 0);
    if (int rc = upload_pulses(h, ampl)) return rc;        // eps <- the amplitudes themselves
    if (grad) {
        CUDA_TRY(h, cudaMemcpyAsync(h->d_damp, dampl, sizeof(double) * LNT, cudaMemcpyHostToDevice, h->stream));
    }
    h->p.shape = nullptr;
    h->p.dshape = grad ? h->d_damp : nullptr;
    run_xchg_begin(h, grad);
    run_formU(h); rec(h, 1);
    run_forward(h, grad, false); rec(h, 2);
    rec(h, 3);
    if (grad) { run_backward(h, nullptr); rec(h, 4); run_gradient(h); }
    else rec(h, 4);
    run_finalize(h, grad); rec(h, 5);
    h->p.shape = shape0;
    h->p.dshape = dshape0;
    if (int rc = download_enqueue(h)) return rc;
    return eval_wait(h, grad);
}

int grape_b200_eval_f_amplitudes(grape_b200_handle* h, const double* ampl, double* J_parts, double* tau) {
    if (!h || !ampl) return GRAPE_B200_EINVAL;
    if (int rc = eval_amplitudes(h, ampl, nullptr, false)) return rc;
    if (h->p.functional == GRAPE_B200_JT_HOST) h->h_out[h->off_J] = 0.0 / 0.0;
    copy_out_common(h, J_parts, tau);
    return check_flags(h);
}

int grape_b200_eval_fg_amplitudes(grape_b200_handle* h, const double* ampl, const double* dampl, double* G_slots,
                                  double* J_parts, double* tau) {
    if (!h || !ampl || !dampl || !G_slots) return GRAPE_B200_EINVAL;
    if (h->p.functional == GRAPE_B200_JT_HOST) {
        h->err = "eval_fg_amplitudes needs a built-in functional";
        return GRAPE_B200_EINVAL;
    }
    if (int rc = eval_amplitudes(h, ampl, dampl, true)) return rc;
    return eval_fg_copy_out(h, G_slots, J_parts, tau, nullptr, nullptr);
}

int grape_b200_forward(grape_b200_handle* h, const double* pulsevals, double* tau, double* sums) {
    if (!h || !pulsevals) return GRAPE_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int64_t l0 = h->launches;
    h->xchg_mode = false;   // split host API: LOCAL partial sums, the caller reduces them over the shards
    rec(h, 0);
    if (int rc = upload_pulses(h, pulsevals)) return rc;
    run_formU(h); rec(h, 1);
    run_forward(h); rec(h, 2); rec(h, 3); rec(h, 4); rec(h, 5);
    // only sums + tau + flags are needed, but one copy of the block is cheapest
    if (int rc = download_all(h)) return rc;
    collect_timings(h, l0);
    h->forward_done = true; h->backward_done = false; h->taugrads_valid = false;
    if (tau) memcpy(tau, h->h_out + h->off_tau, 2 * sizeof(double) * h->p.K);
    if (sums) memcpy(sums, h->h_out + h->off_sums, 4 * sizeof(double));
    return check_flags(h);
}

static int backward_common(grape_b200_handle* h, const cplx* chi_host, double* G_partial,
                           double* J_parts, double* J_b_partial, double* grad_J_a) {
    const int64_t l0 = h->launches;
    h->xchg_mode = false;   // split host API: LOCAL partial gradient
    rec(h, 0); rec(h, 1); rec(h, 2); rec(h, 3);
    run_backward(h, chi_host); rec(h, 4);
    run_gradient(h);
    run_finalize(h, true); rec(h, 5);
    if (int rc = download_all(h)) return rc;
    collect_timings(h, l0);
    h->backward_done = true; h->taugrads_valid = false;
    const int LNT = h->LNT;
    if (G_partial) memcpy(G_partial, h->h_out + LNT, sizeof(double) * LNT);   // grad_J_Tb partial
    if (grad_J_a) memcpy(grad_J_a, h->h_out + 2 * LNT, sizeof(double) * LNT);
    if (J_parts) memcpy(J_parts, h->h_out + h->off_J, 3 * sizeof(double));
    if (J_b_partial) *J_b_partial = h->h_out[h->off_sums + 3];
    return check_flags(h);
}

int grape_b200_backward(grape_b200_handle* h, const double* sums_global, double* G_partial,
                        double* J_parts, double* grad_J_a) {
    if (!h || !sums_global) return GRAPE_B200_EINVAL;
    if (!h->forward_done) { h->err = "grape_b200_backward called before grape_b200_forward"; return GRAPE_B200_ESTATE; }
    if (h->p.functional == GRAPE_B200_JT_HOST) { h->err = "JT_HOST: use grape_b200_backward_chi"; return GRAPE_B200_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    memcpy(h->h_in + h->LNT, sums_global, 4 * sizeof(double));
    CUDA_TRY(h, cudaMemcpyAsync(h->p.sums, h->h_in + h->LNT, 4 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return backward_common(h, nullptr, G_partial, J_parts, nullptr, grad_J_a);
}

int grape_b200_backward_chi(grape_b200_handle* h, const double* chiT, double* G_partial,
                            double* J_b_partial, double* grad_J_a) {
    if (!h || !chiT) return GRAPE_B200_EINVAL;
    if (!h->forward_done) { h->err = "grape_b200_backward_chi called before grape_b200_forward"; return GRAPE_B200_ESTATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_chi_host, chiT, sizeof(cplx) * (size_t)h->p.K * h->p.N,
                                cudaMemcpyHostToDevice, h->stream));
    return backward_common(h, h->d_chi_host, G_partial, nullptr, J_b_partial, grad_J_a);
}

__global__ void __launch_bounds__(256) combine_grad(DevP p) {
    // G = grad_J_Tb + lambda_a * grad_J_a   (after the caller all-reduced grad_J_Tb in place)
    const int LNT = p.L * p.NT;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < LNT; idx += gridDim.x * blockDim.x)
        p.grad[idx] = p.ja_kind ? fma(p.lambda_a, p.grad[2 * LNT + idx], p.grad[LNT + idx]) : p.grad[LNT + idx];
}

int grape_b200_enqueue_forward(grape_b200_handle* h, const double* d_pulsevals) {
    if (!h || !d_pulsevals) return GRAPE_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    // keep a handle-owned copy: lazily evaluated read-backs (fw_storage of the Hermitian schedule) need the pulses later
    if (d_pulsevals != h->d_eps_own)
        CUDA_TRY(h, cudaMemcpyAsync(h->d_eps_own, d_pulsevals, sizeof(double) * h->LNT, cudaMemcpyDeviceToDevice, h->stream));
    h->p.eps = h->d_eps_own;
    h->xchg_mode = true;    // attached peers: sums / gradient are exchanged inside enqueue_forward / enqueue_backward
    rec(h, 0);
    CUDA_TRY(h, cudaMemsetAsync(h->p.flags, 0, sizeof(DevFlags), h->stream));
    run_xchg_begin(h, true);
    run_formU(h); rec(h, 1);
    run_forward(h, true, h->p.functional != GRAPE_B200_JT_HOST); rec(h, 2); rec(h, 3);   // enqueue_backward follows
    h->forward_done = true; h->backward_done = false; h->taugrads_valid = false;
    return 0;
}
int grape_b200_enqueue_backward(grape_b200_handle* h) {
    if (!h) return GRAPE_B200_EINVAL;
    if (!h->forward_done) { h->err = "enqueue_backward before enqueue_forward"; return GRAPE_B200_ESTATE; }
    if (h->p.functional == GRAPE_B200_JT_HOST) { h->err = "JT_HOST: use grape_b200_backward_chi"; return GRAPE_B200_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    run_backward(h, nullptr); rec(h, 4);
    run_gradient(h);
    run_finalize(h, true); rec(h, 5); rec(h, 6);
    h->backward_done = true; h->taugrads_valid = false;
    return 0;
}
int grape_b200_enqueue_combine(grape_b200_handle* h) {
    if (!h) return GRAPE_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int blocks = (h->LNT + 255) / 256;
    combine_grad<<<blocks < 296 ? blocks : 296, 256, 0, h->stream>>>(h->p);
    // J_parts again, from the sums as they are NOW: for functionals whose chi does not couple the trajectories
    // (J_T_re, J_T_ss) the caller may all-reduce sums[4] together with grad_J_Tb, i.e. after enqueue_backward
    finalize_J<<<1, 256, 0, h->stream>>>(h->p, 0);
    h->launches += 2;
    return 0;
}
int grape_b200_finish(grape_b200_handle* h) {
    if (!h) return GRAPE_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_out + h->off_flags, h->p.flags, sizeof(DevFlags), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return check_flags(h);
}

int grape_b200_eval_fg_device(grape_b200_handle* h, const double* d_pulsevals, double* d_G, double* d_J_parts) {
    if (!h || !d_pulsevals) return GRAPE_B200_EINVAL;
    if (h->p.functional == GRAPE_B200_JT_HOST) { h->err = "eval_fg_device needs a built-in functional"; return GRAPE_B200_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int64_t l0 = h->launches;
    rec(h, 0);
    // keep a handle-owned copy: lazily evaluated read-backs (fw_storage of the Hermitian schedule) need the pulses later
    if (d_pulsevals != h->d_eps_own)
        CUDA_TRY(h, cudaMemcpyAsync(h->d_eps_own, d_pulsevals, sizeof(double) * h->LNT, cudaMemcpyDeviceToDevice, h->stream));
    h->p.eps = h->d_eps_own;
    h->xchg_mode = true;
    CUDA_TRY(h, cudaMemsetAsync(h->p.flags, 0, sizeof(DevFlags), h->stream));
    run_xchg_begin(h, true);
    run_formU(h); rec(h, 1);
    run_forward(h, true, true); rec(h, 2); rec(h, 3);
    run_backward(h, nullptr); rec(h, 4);
    run_gradient(h);
    run_finalize(h, true); rec(h, 5);
    if (d_G) CUDA_TRY(h, cudaMemcpyAsync(d_G, h->p.grad, sizeof(double) * h->LNT, cudaMemcpyDeviceToDevice, h->stream));
    if (d_J_parts) CUDA_TRY(h, cudaMemcpyAsync(d_J_parts, h->p.Jparts, 3 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_out + h->off_flags, h->p.flags, sizeof(DevFlags), cudaMemcpyDeviceToHost, h->stream));
    rec(h, 6);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    collect_timings(h, l0);
    h->p.eps = h->d_eps_own;
    h->forward_done = true; h->backward_done = true; h->taugrads_valid = false;
    return check_flags(h);
}

static int ensure_tmp(grape_b200_handle* h, size_t elems) {
    if (h->tmp_elems >= elems) return 0;
    cplx* q;
    if (int rc = dev_alloc(h, &q, elems)) return rc;
    h->d_tmp = q; h->tmp_elems = elems;
    return 0;
}

__global__ void gather_small_states(const cplx* __restrict__ psi, cplx* __restrict__ out, int K, int N, int NT, int k) {
    // out[n*N + i] = psi[n][i][k]
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (NT + 1) * N) out[idx] = psi[(size_t)idx * K + k];
}
__global__ void gather_small_final(const cplx* __restrict__ psi, cplx* __restrict__ out, int K, int N, int NT) {
    // out[k*N + i] = psi[NT][i][k]
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < K * N) {
        const int k = idx / N, i = idx % N;
        out[idx] = psi[((size_t)NT * N + i) * K + k];
    }
}

int grape_b200_get_final_states(grape_b200_handle* h, double* out) {
    if (!h || !out) return GRAPE_B200_EINVAL;
    if (!h->forward_done) { h->err = "no forward sweep has been run"; return GRAPE_B200_ESTATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const DevP& p = h->p;
    const size_t cnt = (size_t)p.K * p.N;
    if (int rc = ensure_tmp(h, cnt)) return rc;
    if (h->path == GRAPE_B200_PATH_SMALL) {
        gather_small_final<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(p.psi, h->d_tmp, p.K, p.N, p.NT);
        h->launches++;
    } else if (h->path == GRAPE_B200_PATH_WARP) {
        warp_gather_final(h->warp, p, h->d_tmp, h->stream, h->launches);
    } else {
        dense_gather_final(h->dense, p, h->d_tmp, h->stream, h->launches);
    }
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_tmp, cnt * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int grape_b200_get_stored_states(grape_b200_handle* h, int32_t k, double* out) {
    if (!h || !out || k < 0 || k >= h->p.K) return GRAPE_B200_EINVAL;
    if (!h->forward_done) { h->err = "no forward sweep has been run"; return GRAPE_B200_ESTATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const DevP& p = h->p;
    const size_t cnt = (size_t)(p.NT + 1) * p.N;
    if (int rc = ensure_tmp(h, cnt)) return rc;
    run_fill_interior(h);
    if (h->path == GRAPE_B200_PATH_SMALL) {
        gather_small_states<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(p.psi, h->d_tmp, p.K, p.N, p.NT, k);
        h->launches++;
    } else if (h->path == GRAPE_B200_PATH_WARP) {
        warp_gather_states(h->warp, p, k, h->d_tmp, h->stream, h->launches);
    } else {
        dense_gather_states(h->dense, p, k, h->d_tmp, h->stream, h->launches);
    }
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_tmp, cnt * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int grape_b200_get_chi_states(grape_b200_handle* h, double* chi, double* norms) {
    if (!h) return GRAPE_B200_EINVAL;
    if (!h->backward_done) { h->err = "no backward sweep has been run"; return GRAPE_B200_ESTATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (chi) CUDA_TRY(h, cudaMemcpyAsync(chi, h->p.chiT, sizeof(cplx) * (size_t)h->p.K * h->p.N, cudaMemcpyDeviceToHost, h->stream));
    if (norms) CUDA_TRY(h, cudaMemcpyAsync(norms, h->p.rho, sizeof(double) * h->p.K, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int grape_b200_get_tau_grads(grape_b200_handle* h, int32_t k, double* out) {
    if (!h || !out || k < 0 || k >= h->p.K) return GRAPE_B200_EINVAL;
    if (!h->backward_done) { h->err = "no backward sweep has been run"; return GRAPE_B200_ESTATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    DevP& p = h->p;
    const size_t per = (size_t)p.L * p.NT;
    if (h->path == GRAPE_B200_PATH_DENSE) {
        h->err = "tau_grads dump is not available on the dense path";
        return GRAPE_B200_EINVAL;
    }
    if (!h->taugrads_valid) {
        if (!h->d_taugrads)
            if (int rc = dev_alloc(h, &h->d_taugrads, per * p.K)) return rc;
        // re-run the contraction of the last backward sweep with the dump enabled (states are still resident);
        // the dump pointer is only set for this launch, so evaluation calls never pay for it
        p.taugrads = h->d_taugrads;
        run_gradient(h);
        p.taugrads = nullptr;
        h->taugrads_valid = true;
    }
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_taugrads + (size_t)k * per, per * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int grape_b200_get_timings(grape_b200_handle* h, double* out8) {
    if (!h || !out8) return GRAPE_B200_EINVAL;
    memcpy(out8, h->timings, sizeof h->timings);
    return 0;
}
int grape_b200_set_profiling(grape_b200_handle* h, int32_t on) {
    if (!h) return GRAPE_B200_EINVAL;
    h->profiling = on != 0;
    return 0;
}
void* grape_b200_device_ptr(grape_b200_handle* h, int32_t which) {
    if (!h) return nullptr;
    switch (which) {
        case 0: return h->p.grad + h->LNT;   // local grad_J_Tb partial
        case 1: return h->p.sums;
        case 2: return h->d_eps_own;
        case 3: return h->p.grad;            // full G
        case 4: return h->p.Jparts;          // J_parts[3]
        case 5: return h->p.tau;             // tau[K] complex (local trajectories)
    }
    return nullptr;
}
void* grape_b200_stream(grape_b200_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t grape_b200_launch_count(const grape_b200_handle* h) { return h ? h->launches : 0; }
int grape_b200_gradient_form(grape_b200_handle* h) {
    if (!h) return -GRAPE_B200_EINVAL;
    if (h->path != GRAPE_B200_PATH_DENSE || !h->dense.kd.on) return 0;
    int ok = 0;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess ||
        cudaMemcpy(&ok, h->dense.kd.ok, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        h->err = "CUDA error while reading the gradient form";
        return -GRAPE_B200_ECUDA;
    }
    return ok ? (h->dense.d.nstrip > 1 ? 2 : 1) : 0;
}

int grape_b200_dense_concurrent(grape_b200_handle* h) {
    if (!h) return -GRAPE_B200_EINVAL;
    if (h->path != GRAPE_B200_PATH_DENSE || !h->dense.concurrent || !h->dense.dual_launched) return 0;
    int ok = 0;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess ||
        cudaMemcpy(&ok, h->dense.kd.ok, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        h->err = "CUDA error while reading the gradient form";
        return -GRAPE_B200_ECUDA;
    }
    return ok ? 1 : 0;
}

int grape_b200_dense_orders(grape_b200_handle* h, int* orders) {
    if (!h || !orders) return -GRAPE_B200_EINVAL;
    if (h->path != GRAPE_B200_PATH_DENSE || !h->dense.kd.on) { if (h) h->err = "no Krylov-form schedule on this handle"; return -GRAPE_B200_EINVAL; }
    if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess ||
        cudaMemcpy(orders, h->dense.kd.m_n, sizeof(int) * (size_t)h->p.NT, cudaMemcpyDeviceToHost) != cudaSuccess) {
        h->err = "CUDA error while reading the step orders";
        return -GRAPE_B200_ECUDA;
    }
    return h->dense.d.econ ? 1 : 0;
}

int grape_b200_econ_table(int m, double* theta, double* g) {
    if (m < 2 || m > ECON_MAXM || !theta || !g) return GRAPE_B200_EINVAL;
    const EconTab& t = econ_table();
    *theta = t.theta[m];
    for (int j = 0; j <= m; ++j) g[j] = t.g[m][j];
    return 0;
}

int grape_b200_small_schedule(grape_b200_handle* h) {
    if (!h) return -GRAPE_B200_EINVAL;
    if (h->path != GRAPE_B200_PATH_SMALL || !h->seg_on) return 0;
    if (!h->seg_herm) return 1;
    if (!seg_real_active(h)) return 2;
    if (h->sym_v != 1) return 3;   // the staged kernels serve every step themselves
    int nf = 0;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess ||
        cudaMemcpy(&nf, h->seg.notfast, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        h->err = "CUDA error while reading the schedule flag";
        return -GRAPE_B200_ECUDA;
    }
    return nf ? 2 : 3;
}


// ------------------------------------------------------------------------------------------------
// Peer exchange over NVLink (xchg.cuh): multi-GPU trajectory sharding behind the ABI
// ------------------------------------------------------------------------------------------------
namespace {
struct XchgLayout { size_t flag_off, slot_off, epoch_off, bytes; int XS; };
XchgLayout xchg_layout(int world, int LNT) {
    XchgLayout l;
    l.XS = (LNT + 4 + 1) & ~1;
    l.flag_off = 0;
    size_t o = (size_t)2 * world * XCHG_MAXB * sizeof(unsigned long long);
    o = (o + 255) & ~(size_t)255;
    l.slot_off = o;
    o += (size_t)2 * 2 * world * l.XS * sizeof(double);
    o = (o + 255) & ~(size_t)255;
    l.epoch_off = o;
    l.bytes = o + 256;
    return l;
}
void xchg_release(H* h) {
    for (void* q : h->xchg_ipc_opened) cudaIpcCloseMemHandle(q);
    h->xchg_ipc_opened.clear();
    if (h->xchg_buf) { cudaFree(h->xchg_buf); h->xchg_buf = nullptr; }
    h->xchg_on = false;
    memset(&h->xd, 0, sizeof h->xd);
    drop_graphs(h);
}
void xchg_point(H* h, int r, void* base) {
    const XchgLayout l = xchg_layout(h->xd.world, h->LNT);
    h->xd.flags[r] = reinterpret_cast<unsigned long long*>(static_cast<char*>(base) + l.flag_off);
    h->xd.slots[r] = reinterpret_cast<double*>(static_cast<char*>(base) + l.slot_off);
}
int xchg_init_impl(H* h, int rank, int world) {
    if (world < 1 || world > XCHG_MAXW || rank < 0 || rank >= world) {
        h->err = "xchg_init: need 0 <= rank < world <= 16";
        return GRAPE_B200_EINVAL;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    xchg_release(h);
    const XchgLayout l = xchg_layout(world, h->LNT);
    CUDA_TRY(h, cudaMalloc(&h->xchg_buf, l.bytes));
    CUDA_TRY(h, cudaMemset(h->xchg_buf, 0, l.bytes));
    h->xchg_bytes = l.bytes;
    h->xd.rank = rank; h->xd.world = world; h->xd.XS = l.XS;
    h->xd.epoch = reinterpret_cast<unsigned long long*>(static_cast<char*>(h->xchg_buf) + l.epoch_off);
    h->xd.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(h->xchg_buf) + l.epoch_off + 64);
    h->xd.timeout = &h->p.flags->xchg_timeout;
    xchg_point(h, rank, h->xchg_buf);
    return 0;
}
}  // namespace

int grape_b200_xchg_init(grape_b200_handle* h, int32_t rank, int32_t world, void* ipc_handle_out) {
    if (!h) return GRAPE_B200_EINVAL;
    if (int rc = xchg_init_impl(h, rank, world)) return rc;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t mh;
        CUDA_TRY(h, cudaIpcGetMemHandle(&mh, h->xchg_buf));
        static_assert(sizeof(mh) == GRAPE_B200_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
        memcpy(ipc_handle_out, &mh, sizeof mh);
    }
    return 0;
}

int grape_b200_xchg_attach(grape_b200_handle* h, const void* ipc_handles) {
    if (!h || !ipc_handles) return GRAPE_B200_EINVAL;
    if (!h->xchg_buf) { h->err = "xchg_attach before xchg_init"; return GRAPE_B200_ESTATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    for (int r = 0; r < h->xd.world; ++r) {
        if (r == h->xd.rank) continue;
        cudaIpcMemHandle_t mh;
        memcpy(&mh, static_cast<const char*>(ipc_handles) + (size_t)r * GRAPE_B200_IPC_HANDLE_BYTES, sizeof mh);
        void* base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            char b[256];
            snprintf(b, sizeof b, "cudaIpcOpenMemHandle of rank %d's exchange buffer failed: %s (peers need NVLink/P2P access "
                     "and a shared IPC namespace)", r, cudaGetErrorString(e));
            h->err = b;
            cudaGetLastError();
            return GRAPE_B200_ENCCL;
        }
        h->xchg_ipc_opened.push_back(base);
        xchg_point(h, r, base);
    }
    h->xchg_on = h->xd.world > 1;
    drop_graphs(h);
    return 0;
}

int grape_b200_xchg_detach(grape_b200_handle* h) {
    if (!h) return GRAPE_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    xchg_release(h);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// One host process driving several GPUs (the `ccall` host of INTEGRATION.md)
// ------------------------------------------------------------------------------------------------
struct grape_b200_multi {
    std::vector<grape_b200_handle*> sh;
    std::vector<int> k_lo;
    int K, N, LNT;
    std::string err;
};
namespace { thread_local std::string g_multi_error; }

const char* grape_b200_multi_last_error(const grape_b200_multi* m) { return m ? m->err.c_str() : g_multi_error.c_str(); }

void grape_b200_multi_destroy(grape_b200_multi* m) {
    if (!m) return;
    for (auto* h : m->sh) if (h) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); }
    for (auto* h : m->sh) grape_b200_destroy(h);
    delete m;
}

int grape_b200_multi_create(const grape_b200_problem* d, const int32_t* devices, int32_t ndev, grape_b200_multi** out) {
    if (!out) return GRAPE_B200_EINVAL;
    *out = nullptr;
    auto fail = [&](int code, const std::string& msg) { g_multi_error = msg; return code; };
    if (!d || !devices || ndev < 1 || ndev > XCHG_MAXW) return fail(GRAPE_B200_EINVAL, "multi_create: need 1 <= ndev <= 16 devices");
    if (d->K < ndev) return fail(GRAPE_B200_EINVAL, "multi_create: fewer trajectories than devices");
    if (d->K <= 0 || d->N <= 0 || d->L <= 0 || d->NT <= 0 || d->G <= 0 || !d->H0 || !d->Hc || !d->psi0 || !d->tgt)
        return fail(d->L <= 0 ? GRAPE_B200_ENOCONTROLS : GRAPE_B200_EINVAL,
                    d->L <= 0 ? "no controls in trajectories: cannot optimize" : "multi_create: invalid descriptor");
    if (!d->gen_of_traj && d->G != 1 && d->G != d->K)
        return fail(GRAPE_B200_EINVAL, "gen_of_traj is required unless G == 1 or G == K");
    grape_b200_multi* m = new grape_b200_multi();
    m->K = d->K; m->N = d->N; m->LNT = d->L * d->NT;
    const size_t NN2 = (size_t)2 * d->N * d->N, N2 = (size_t)2 * d->N;
    for (int i = 0; i < ndev; ++i) {
        // contiguous block of trajectories; generators no local trajectory uses are dropped
        const int lo = (int)(((long long)d->K * i) / ndev), hi = (int)(((long long)d->K * (i + 1)) / ndev);
        const int Kl = hi - lo;
        std::vector<int> gens, gl(Kl);
        for (int k = lo; k < hi; ++k) gens.push_back(d->gen_of_traj ? d->gen_of_traj[k] : (d->G == 1 ? 0 : k));
        std::vector<int> uniq(gens);
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        for (int k = 0; k < Kl; ++k) gl[k] = (int)(std::lower_bound(uniq.begin(), uniq.end(), gens[k]) - uniq.begin());
        const int Gl = (int)uniq.size();
        std::vector<double> H0((size_t)Gl * NN2), Hc((size_t)Gl * d->L * NN2);
        for (int g = 0; g < Gl; ++g) {
            if (uniq[g] < 0 || uniq[g] >= d->G) { grape_b200_multi_destroy(m); return fail(GRAPE_B200_EINVAL, "gen_of_traj entry out of range"); }
            memcpy(&H0[(size_t)g * NN2], d->H0 + (size_t)uniq[g] * NN2, NN2 * sizeof(double));
            memcpy(&Hc[(size_t)g * d->L * NN2], d->Hc + (size_t)uniq[g] * d->L * NN2, (size_t)d->L * NN2 * sizeof(double));
        }
        grape_b200_problem ld = *d;
        ld.K = Kl; ld.G = Gl; ld.K_global = d->K_global > 0 ? d->K_global : d->K; ld.device = devices[i];
        ld.gen_of_traj = gl.data(); ld.H0 = H0.data(); ld.Hc = Hc.data();
        ld.psi0 = d->psi0 + (size_t)lo * N2; ld.tgt = d->tgt + (size_t)lo * N2;
        ld.weights = d->weights ? d->weights + lo : nullptr;
        if (d->gb_kind != GRAPE_B200_GB_NONE && d->gb_D && d->gb_nD == d->K) { ld.gb_D = d->gb_D + (size_t)lo * NN2; ld.gb_nD = Kl; }
        grape_b200_handle* h = nullptr;
        const int rc = grape_b200_create(&ld, &h);
        if (rc) { g_multi_error = g_create_error; grape_b200_multi_destroy(m); return rc; }
        m->sh.push_back(h);
        m->k_lo.push_back(lo);
    }
    if (ndev > 1) {
        // every device maps every other device's exchange buffer (NVLink / NVSwitch peer access)
        for (int i = 0; i < ndev; ++i) {
            cudaSetDevice(m->sh[i]->device);
            for (int j = 0; j < ndev; ++j) {
                if (i == j || m->sh[i]->device == m->sh[j]->device) continue;   // same device: plain device pointers
                int can = 0;
                cudaDeviceCanAccessPeer(&can, m->sh[i]->device, m->sh[j]->device);
                if (!can) { grape_b200_multi_destroy(m); return fail(GRAPE_B200_ENCCL, "multi_create: devices have no peer (NVLink/P2P) access to each other"); }
                cudaError_t e = cudaDeviceEnablePeerAccess(m->sh[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    grape_b200_multi_destroy(m);
                    return fail(GRAPE_B200_ENCCL, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                }
                cudaGetLastError();
            }
        }
        for (int i = 0; i < ndev; ++i)
            if (int rc = xchg_init_impl(m->sh[i], i, ndev)) { g_multi_error = m->sh[i]->err; grape_b200_multi_destroy(m); return rc; }
        for (int i = 0; i < ndev; ++i) {
            for (int j = 0; j < ndev; ++j) if (j != i) xchg_point(m->sh[i], j, m->sh[j]->xchg_buf);
            m->sh[i]->xchg_on = true;
        }
    }
    *out = m;
    return GRAPE_B200_OK;
}

int32_t grape_b200_multi_size(const grape_b200_multi* m) { return m ? (int32_t)m->sh.size() : 0; }
grape_b200_handle* grape_b200_multi_shard(grape_b200_multi* m, int32_t i, int32_t* k_first) {
    if (!m || i < 0 || i >= (int)m->sh.size()) return nullptr;
    if (k_first) *k_first = m->k_lo[i];
    return m->sh[i];
}

static int multi_eval(grape_b200_multi* m, const double* pulsevals, bool grad, double* G, double* J_parts, double* tau,
                      double* grad_J_Tb, double* grad_J_a) {
    if (!m || !pulsevals || (grad && !G)) return GRAPE_B200_EINVAL;
    // every shard is enqueued before any is waited for: the shards meet in the exchange kernels
    int rc = 0;
    size_t launched = 0;
    for (; launched < m->sh.size(); ++launched) {
        grape_b200_handle* h = m->sh[launched];
        if (grad && h->p.functional == GRAPE_B200_JT_HOST) { h->err = "multi_eval_fg needs a built-in functional"; rc = GRAPE_B200_EINVAL; }
        else rc = eval_launch(h, pulsevals, grad);
        if (rc) { m->err = h->err; break; }
    }
    for (size_t i = 0; i < launched; ++i) {
        const int rw = eval_wait(m->sh[i], grad);
        if (rw && !rc) { rc = rw; m->err = m->sh[i]->err; }
    }
    if (rc) return rc;
    for (size_t i = 0; i < m->sh.size(); ++i) {
        grape_b200_handle* h = m->sh[i];
        const int rf = check_flags(h);
        if (rf && !rc) { rc = rf; m->err = h->err; }
        if (tau) memcpy(tau + (size_t)2 * m->k_lo[i], h->h_out + h->off_tau, 2 * sizeof(double) * h->p.K);
    }
    grape_b200_handle* h0 = m->sh[0];
    if (grad) {
        memcpy(G, h0->h_out, sizeof(double) * m->LNT);
        if (grad_J_Tb) memcpy(grad_J_Tb, h0->h_out + m->LNT, sizeof(double) * m->LNT);
        if (grad_J_a) memcpy(grad_J_a, h0->h_out + 2 * (size_t)m->LNT, sizeof(double) * m->LNT);
    }
    if (J_parts) memcpy(J_parts, h0->h_out + h0->off_J, 3 * sizeof(double));
    return rc;
}
int grape_b200_multi_eval_f(grape_b200_multi* m, const double* pulsevals, double* J_parts, double* tau) {
    return multi_eval(m, pulsevals, false, nullptr, J_parts, tau, nullptr, nullptr);
}
int grape_b200_multi_eval_fg(grape_b200_multi* m, const double* pulsevals, double* G, double* J_parts, double* tau,
                             double* grad_J_Tb, double* grad_J_a) {
    return multi_eval(m, pulsevals, true, G, J_parts, tau, grad_J_Tb, grad_J_a);
}
int grape_b200_multi_get_final_states(grape_b200_multi* m, double* out) {
    if (!m || !out) return GRAPE_B200_EINVAL;
    for (size_t i = 0; i < m->sh.size(); ++i) {
        const int rc = grape_b200_get_final_states(m->sh[i], out + (size_t)2 * m->N * m->k_lo[i]);
        if (rc) { m->err = m->sh[i]->err; return rc; }
    }
    return 0;
}

}  // extern "C"
