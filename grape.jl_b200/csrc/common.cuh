// common.cuh -- shared device helpers for the GRAPE B200 engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;

#define GB_HD __host__ __device__ __forceinline__
#define GB_D __device__ __forceinline__

GB_HD cplx mk(double re, double im) { cplx r; r.x = re; r.y = im; return r; }
GB_D cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
GB_D cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
GB_D cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }
GB_D cplx cmul(cplx a, cplx b) {
    return mk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// acc += a*b
GB_D void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
GB_D void cfmac(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.y, b.x, acc.y);
}
// acc += s*b  (real s)
GB_D void cfmar(cplx& acc, double s, cplx b) {
    acc.x = fma(s, b.x, acc.x);
    acc.y = fma(s, b.y, acc.y);
}
GB_D double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }
GB_D double cnorm2(cplx a) { return fma(a.x, a.x, a.y * a.y); }

// cp.async (LDGSTS) 16-byte, L2-only caching: streams that are read exactly once
GB_D void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
// same with zero fill: copies 16 bytes if `valid`, else writes zeros (src is not read)
GB_D void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(n) : "memory");
}
GB_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NPEND>
GB_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NPEND) : "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on a shared-memory mbarrier: one elected thread
// hands whole rows to the copy engine instead of every thread issuing 16-byte LDGSTS.
GB_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
GB_D void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// orders the mbarrier initialisation (generic proxy) before its use by the async proxy
GB_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// generic-proxy accesses to shared memory (the warps' reads of the old tile) before async-proxy writes to it
GB_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
GB_D void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
GB_D void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
GB_D void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// streaming (evict-first) 16-byte store for write-once storage
GB_D void st_cs(cplx* p, cplx v) { __stcs(p, v); }
GB_D cplx ld_cs(const cplx* p) { return __ldcs(p); }

// 1/j!  j = 0..20
__constant__ double c_invfact[21] = {
    1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
    1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0,
    1.0 / 87178291200.0, 1.0 / 1307674368000.0, 1.0 / 20922789888000.0,
    1.0 / 355687428096000.0, 1.0 / 6402373705728000.0, 1.0 / 121645100408832000.0,
    1.0 / 2432902008176640000.0};

// Matrix-exponential plan: Taylor degree class (3, 7, 11, 15) and number of
// squarings s for a matrix of 1-norm `nrm`.  Truncation bound
// theta^(d+1)/(d+1)! <= 1.1e-16 (relative to ||exp|| ~ 1):
//   d=3: 2.2e-4   d=7: 3.8e-2   d=11: 0.2476   d=15: 0.684   (a safety margin is applied)
GB_HD void exp_plan(double nrm, int& degree, int& s) {
    s = 0;
    if (nrm <= 2.0e-4) { degree = 3; return; }
    if (nrm <= 3.5e-2) { degree = 7; return; }
    if (nrm <= 0.23) { degree = 11; return; }
    degree = 15;
    while (nrm > 0.65 && s < 60) { nrm *= 0.5; ++s; }
}

// Number of Taylor terms m for the vector-form block recursion so that
// theta^m/m! <= 2e-17, and number of sub-steps 2^s with theta <= 1.
GB_HD void vec_plan(double nrm, int& m, int& s) {
    s = 0;
    while (nrm > 1.0 && s < 30) { nrm *= 0.5; ++s; }
    // theta^m/m! <= 2e-17  <=>  theta^m <= 2e-17 * m!   (no divisions: this runs once per unit)
    double t = nrm, f = 1.0;
    m = 1;
    while (t > 2e-17 * f && m < 40) { ++m; t *= nrm; f *= (double)m; }
    if (m < 2) m = 2;
}

// flags written by kernels, read by the host after the final sync
struct DevFlags {
    int chi_bad_k;        // 1-based index of a trajectory with rho < chi_min_norm, 0 = ok
    int taylor_fail;      // != 0: taylor_grad_step did not converge
    double chi_bad_rho;
    double taylor_r;
    int xchg_timeout;     // != 0: a peer never arrived at an exchange (xchg.cuh)
    int pad_;
};

// Everything a kernel needs, passed by value.
struct DevP {
    int K, N, L, NT, G, Kglobal;
    int functional, grad_method, ja_kind, gb_kind, gb_nD;
    int taylor_max_order, taylor_check;
    double lambda_a, lambda_b, chi_min_norm, taylor_tol;
    const double* tlist;   // [NT+1]
    const double* eps;     // [L*NT] pulse values (device copy of `pulsevals`)
    const double* shape;   // [L*NT] or nullptr: the amplitude of control l at step n is eps * shape
    const double* dshape;  // [L*NT] or nullptr (= 1): d amplitude / d eps; == shape for ShapedAmplitude, the host's
                           // d a / d eps in amplitude mode (grape_b200_eval_fg_amplitudes: eps holds a itself, shape is off)
    const int* gen;        // [K]
    const cplx* H0;        // path-specific layout
    const cplx* Hc;
    const cplx* psi0;
    const cplx* tgt;
    const cplx* D;
    const double* w;       // [K] weights or nullptr
    cplx* U;               // propagators
    cplx* psi;             // fw_storage
    cplx* chi;             // backward states
    cplx* tau;             // [K]
    cplx* chiT;            // [K*N] boundary chi (normalised), AoS [k][i]
    double* rho;           // [K]
    double* jb;            // [K] J_b_trajectory
    double* sums;          // [4]
    double* partial;       // [KB][L*NT]
    double* grad;          // [3][L*NT]: G, grad_J_Tb, grad_J_a
    double* Jparts;        // [3]
    cplx* taugrads;        // optional [K][L][NT] dump (nullptr = off)
    DevFlags* flags;
    int KB;
    const int* KBdev;      // optional device override of KB (dense path: 1 when the Krylov-form contraction ran)
};
