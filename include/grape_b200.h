/*
 * grape_b200.h -- C-ABI of the B200-native GRAPE gradient engine.
 *
 * Drop-in boundary for the gradient hot path of GRAPE.jl (reference tree at
 * /root/reference, citations below are relative to it).  The reference's
 * optimizer back-ends only ever call the closure `fg!(F, G, pulsevals)`
 * (src/optimize.jl:98-111; invoked at ext/GRAPELBFGSBExt.jl:99 and through
 * Optim at ext/GRAPEOptimExt.jl:31).  `fg!` is `evaluate_functional`
 * (src/optimize.jl:696-768) when no gradient is wanted and
 * `evaluate_gradient!` (src/optimize.jl:824-1014) otherwise, both operating on
 * the `GrapeWrk` workspace (src/workspace.jl:78-362).  This library replaces
 * exactly those three things:
 *
 *   GrapeWrk(trajectories, tlist, kwargs)   -> grape_b200_create
 *   evaluate_functional(pulsevals, wrk)     -> grape_b200_eval_f
 *   evaluate_gradient!(G, pulsevals, wrk)   -> grape_b200_eval_fg
 *
 * plus a split form (forward / backward) used when trajectories are sharded
 * over several processes (one per GPU) or when J_T / chi are arbitrary host
 * closures, and accessors for the workspace fields that callbacks read
 * (src/optimize.jl:185-216, 402-478).
 *
 * Plain pointers and sizes only; all arrays are caller-owned host memory that
 * is never retained past the call.  Complex numbers are interleaved
 * (re, im) doubles = Julia `ComplexF64`.  Matrices are column-major N x N
 * (Julia layout).  All functions return 0 on success or a GRAPE_B200_E* code;
 * `grape_b200_last_error` returns the message (the Julia shim turns it into
 * `error(msg)` so that the catch at src/optimize.jl:125-135 keeps working).
 *
 * One handle = one CUDA device = one shard of trajectories.  A handle is not
 * re-entrant; distinct handles are independent.
 *
 * Several GPUs (the reference parallelises the same axis with threads: `@threadsif`
 * over k, src/optimize.jl:720, 876): contiguous blocks of trajectories, one shard per
 * GPU.  The only couplings -- the functional / chi (all tau_k, src/optimize.jl:755-760,
 * 845-855) and the sum over k (src/optimize.jl:574-584) -- are reduced BY THE LIBRARY'S
 * OWN KERNELS over NVLink peer memory (csrc/xchg.cuh), either
 *   - in one host process:   grape_b200_multi_create(desc, devices, ndev) + grape_b200_multi_eval_fg
 *     (what a Julia host calls), or
 *   - one process per GPU:   grape_b200_create(local shard) on every rank, grape_b200_xchg_init ->
 *     exchange the 64-byte IPC handles by any means -> grape_b200_xchg_attach; then eval_f / eval_fg /
 *     eval_fg_device / enqueue_* return the GLOBAL J and gradient on every rank (collective calls:
 *     every rank makes the same sequence).
 */
#ifndef GRAPE_B200_H
#define GRAPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRAPE_B200_ABI_VERSION 1

/* error codes */
#define GRAPE_B200_OK            0
#define GRAPE_B200_EINVAL        1  /* bad descriptor / argument                          */
#define GRAPE_B200_ECUDA         2  /* CUDA runtime failure (no device, OOM, launch, ...) */
#define GRAPE_B200_ECHINORM      3  /* chi norm < chi_min_norm (src/optimize.jl:1021-1025) */
#define GRAPE_B200_ETAYLOR       4  /* taylor_grad_step! did not converge (src/optimize.jl:644-648) */
#define GRAPE_B200_ENOCONTROLS   5  /* "no controls in trajectories" (src/workspace.jl:155-157) */
#define GRAPE_B200_ESTATE        6  /* call sequence error (backward before forward, ...)  */
#define GRAPE_B200_ENCCL         7  /* multi-GPU exchange failure (no peer access, IPC, a shard never arrived) */

/* J_T / chi kind: QuantumControl.Functionals J_T_sm / J_T_re / J_T_ss with their
 * analytic chi (make_chi, src/workspace.jl:306-308); HOST = arbitrary closures,
 * served by grape_b200_forward -> host J_T/chi -> grape_b200_backward_chi. */
#define GRAPE_B200_JT_SM   0
#define GRAPE_B200_JT_RE   1
#define GRAPE_B200_JT_SS   2
#define GRAPE_B200_JT_HOST 3

/* gradient_method kwarg (src/docstring.jl:108-128) */
#define GRAPE_B200_GRADGEN 0
#define GRAPE_B200_TAYLOR  1

/* J_a kind: built-in fluence (J_a_fluence); any other J_a stays a host closure
 * because it only touches `pulsevals` (src/optimize.jl:761-763, 1004-1011). */
#define GRAPE_B200_JA_NONE    0
#define GRAPE_B200_JA_FLUENCE 1

/* g_b kind: built-in quadratic form g_b = <Psi|D|Psi>, xi = -D Psi
 * (src/optimize.jl:727-750, 856-866, 897-908; test/test_state_running_cost.jl:17-65) */
#define GRAPE_B200_GB_NONE     0
#define GRAPE_B200_GB_QUADFORM 1

/* execution path selector (0 = choose from N) */
#define GRAPE_B200_PATH_AUTO   0
#define GRAPE_B200_PATH_SMALL  1  /* N <= 4: one thread per (trajectory, step), registers   */
#define GRAPE_B200_PATH_WARP   2  /* N <= 32: sub-warp per (trajectory, step), shuffles     */
#define GRAPE_B200_PATH_DENSE  3  /* any N: polynomial apply on state blocks, DMMA ZGEMM    */
#define GRAPE_B200_PATH_SMALL_CHAIN 4 /* N <= 4 with plain NT-step chains (no time segmentation) */
#define GRAPE_B200_PATH_WARP_CHAIN  5 /* N <= 32 with plain NT-step chains (no time segmentation) */

/* Problem descriptor = the hot fields of GrapeWrk (src/workspace.jl:78-144). */
typedef struct grape_b200_problem {
    int32_t abi_version;      /* GRAPE_B200_ABI_VERSION                                        */
    int32_t K;                /* trajectories held by THIS handle (local shard)                */
    int32_t N;                /* Hilbert-space dimension                                       */
    int32_t L;                /* number of controls                                            */
    int32_t NT;               /* number of time intervals = length(tlist) - 1                  */
    int32_t G;                /* number of distinct generators (1 <= G <= K)                   */
    int32_t K_global;         /* trajectories over all shards (0 -> K); functionals use this   */
    int32_t device;           /* CUDA device ordinal                                           */
    const double*  tlist;     /* [NT+1]                                                        */
    const int32_t* gen_of_traj; /* [K] generator index per trajectory; NULL: 0 if G==1, k if G==K */
    const double*  H0;        /* [G][N*N] complex, column-major: drift of generator g          */
    const double*  Hc;        /* [G][L][N*N] complex, column-major: control operators          */
    const double*  shape;     /* [L][NT] amplitude shape S (ShapedAmplitude) or NULL (= 1)     */
    const double*  psi0;      /* [K][N] complex initial states                                 */
    const double*  tgt;       /* [K][N] complex target states                                  */
    const double*  weights;   /* [K] trajectory weights or NULL (= 1)                          */
    int32_t functional;       /* GRAPE_B200_JT_*                                               */
    int32_t gradient_method;  /* GRAPE_B200_GRADGEN / _TAYLOR                                  */
    int32_t ja_kind;          /* GRAPE_B200_JA_*                                               */
    int32_t gb_kind;          /* GRAPE_B200_GB_*                                               */
    double  lambda_a;         /* src/docstring.jl:96                                           */
    double  lambda_b;         /* src/docstring.jl:104                                          */
    const double* gb_D;       /* [gb_nD][N*N] complex column-major Hermitian D, or NULL        */
    int32_t gb_nD;            /* 1 (shared) or K (one per trajectory)                          */
    int32_t taylor_max_order;         /* taylor_grad_max_order (default 100)                   */
    double  taylor_tolerance;         /* taylor_grad_tolerance (default 1e-16)                 */
    int32_t taylor_check_convergence; /* taylor_grad_check_convergence (default 1)             */
    int32_t path;             /* GRAPE_B200_PATH_*                                             */
    double  chi_min_norm;     /* default 1e-100 (src/optimize.jl:846)                          */
} grape_b200_problem;

typedef struct grape_b200_handle grape_b200_handle;

int  grape_b200_abi_version(void);

/* GrapeWrk constructor (src/workspace.jl:147-362): uploads the static problem once. */
int  grape_b200_create(const grape_b200_problem* desc, grape_b200_handle** out);
void grape_b200_destroy(grape_b200_handle* h);
/* h == NULL -> message of the last failed grape_b200_create on this thread */
const char* grape_b200_last_error(const grape_b200_handle* h);

/* evaluate_functional (src/optimize.jl:696-768).
 * pulsevals [L*NT] blocked by control (src/workspace.jl:159-162);
 * J_parts[3] = (J_T, lambda_a*J_a, lambda_b*J_b)  (src/optimize.jl:755-766);
 * tau [2*K] complex overlaps (src/optimize.jl:753), may be NULL. */
int grape_b200_eval_f(grape_b200_handle* h, const double* pulsevals,
                      double* J_parts, double* tau);

/* evaluate_gradient! (src/optimize.jl:824-1014).
 * G [L*NT] full gradient; grad_J_Tb / grad_J_a [L*NT] may be NULL
 * (wrk.grad_J_Tb, wrk.grad_J_a, src/optimize.jl:1002-1011). */
int grape_b200_eval_fg(grape_b200_handle* h, const double* pulsevals, double* G,
                       double* J_parts, double* tau, double* grad_J_Tb, double* grad_J_a);

/* Amplitude mode: non-linear controls and per-term amplitudes.  The reference differentiates the generator with
 * respect to each control through `get_control_derivs` (src/workspace.jl:283-285) and evaluates mu = dH/d eps per step
 * when it is not a constant operator (src/optimize.jl:946-951; `isnothing(mu)` -> zero gradient).  Here the descriptor's
 * L "controls" are amplitude SLOTS -- one per (control, amplitude) pair the generators contain -- and the host, which
 * owns the closures, passes per call
 *     ampl[i*NT + n]  = a_i(eps_{c(i),n}, t_n)             H_n = H0 + sum_i ampl[i][n] Hc_i
 *     dampl[i*NT + n] = d a_i / d eps_{c(i),n}             mu_{i,n} = dampl[i][n] Hc_i   (0: no dependence)
 * (L*NT scalar evaluations) and receives G_slots[i*NT + n] = dJ_T/d eps_{c(i),n} through slot i; it adds the slots of
 * one control and J_a / grad_J_a itself (they act on the control values: create the handle with GRAPE_B200_JA_NONE).
 * The descriptor's `shape` must be NULL for such a handle.  A linear control is the special case ampl = eps, dampl = 1,
 * a ShapedAmplitude ampl = S eps, dampl = S. */
int grape_b200_eval_f_amplitudes(grape_b200_handle* h, const double* ampl, double* J_parts, double* tau);
int grape_b200_eval_fg_amplitudes(grape_b200_handle* h, const double* ampl, const double* dampl, double* G_slots,
                                  double* J_parts, double* tau);

/* Split form.  forward = src/optimize.jl:696-753 on the local shard.
 * sums[4] = local partial (Re sum_k w_k tau_k, Im sum_k w_k tau_k,
 *            sum_k w_k |tau_k|^2, sum_k J_b_trajectory[k]).
 * The caller all-reduces `sums` over shards (or uses them as they are for one
 * shard) and passes the global values to grape_b200_backward, which evaluates
 * chi (src/optimize.jl:845-869), the backward sweep (:873-994) and the
 * k-reduction (:574-584) for the local trajectories.  G_partial [L*NT] is the
 * LOCAL partial of grad_J_Tb; the caller sums it over shards and adds
 * lambda_a*grad_J_a (returned in full by every shard). */
int grape_b200_forward(grape_b200_handle* h, const double* pulsevals,
                       double* tau, double* sums);
int grape_b200_backward(grape_b200_handle* h, const double* sums_global,
                        double* G_partial, double* J_parts, double* grad_J_a);
/* HOST functional: caller supplies chi_k(T) = -dJ_T/d<Psi_k(T)|  [K][N] complex,
 * un-normalised (src/optimize.jl:845-855); J_T itself stays on the host. */
int grape_b200_backward_chi(grape_b200_handle* h, const double* chiT,
                            double* G_partial, double* J_b_partial, double* grad_J_a);

/* Workspace read-backs for callbacks / result bookkeeping. */
int grape_b200_get_final_states(grape_b200_handle* h, double* out /* [K][N] complex */);   /* fw_propagators[k].state, src/optimize.jl:187-189 */
int grape_b200_get_stored_states(grape_b200_handle* h, int32_t k, double* out /* N x (NT+1) complex col-major */); /* wrk.fw_storage[k], src/workspace.jl:215 */
int grape_b200_get_chi_states(grape_b200_handle* h, double* chi /* [K][N] complex */, double* norms /* [K] */); /* wrk.chi_states, chi_states_norm src/optimize.jl:867-869 */
int grape_b200_get_tau_grads(grape_b200_handle* h, int32_t k, double* out /* NT x L complex col-major */);   /* wrk.tau_grads[k], src/workspace.jl:236-237 */

/* Per-phase device timings of the last eval call, milliseconds (CUDA events):
 * out[0]=H2D+propagator formation, [1]=forward sweep, [2]=tau/chi, [3]=backward sweep,
 * [4]=gradient contraction+reduction, [5]=D2H, [6]=total device, [7]=kernel launches in the call */
int grape_b200_get_timings(grape_b200_handle* h, double* out8);
/* Enable (1) / disable (0) per-phase event timing (adds event records, default off). */
int grape_b200_set_profiling(grape_b200_handle* h, int32_t on);

/* Device-resident entry points for benchmarks / pipelines: same as eval_fg but
 * pulsevals and G are DEVICE pointers on the handle's device; nothing is copied
 * to the host except the error flags.  `stream` is a cudaStream_t (0 = the
 * handle's own stream). */
int grape_b200_eval_fg_device(grape_b200_handle* h, const double* d_pulsevals, double* d_G,
                              double* d_J_parts /* 3, device, may be NULL */);
/* Asynchronous device pipeline (no host synchronisation, nothing copied): used when
 * the two exchanges of a sharded evaluation are done in place on device buffers by
 * the caller's collective library on the handle's stream:
 *   enqueue_forward(d_pulsevals)  -> local sums at device_ptr(1)      [all-reduce in place]
 *   enqueue_backward()            -> local grad_J_Tb at device_ptr(0) [all-reduce in place]
 *   enqueue_combine()             -> G = grad_J_Tb + lambda_a*grad_J_a at device_ptr(3), J_parts from the current sums
 * For J_T_re / J_T_ss (chi_k does not depend on the other trajectories) the first all-reduce may be deferred and
 * fused with the second: enqueue_forward, enqueue_backward, ONE all-reduce of {sums[4], grad_J_Tb}, enqueue_combine.
 *   finish()                      -> synchronise, report chi-norm / Taylor errors
 * With one shard, enqueue_forward + enqueue_backward + finish is a complete
 * evaluate_gradient! (G at device_ptr(3)). */
int grape_b200_enqueue_forward(grape_b200_handle* h, const double* d_pulsevals);
int grape_b200_enqueue_backward(grape_b200_handle* h);
int grape_b200_enqueue_combine(grape_b200_handle* h);
int grape_b200_finish(grape_b200_handle* h);
/* Device pointer of the handle's gradient buffer [L*NT] and sums buffer [4]
 * (for in-place collectives by the caller). */
void* grape_b200_device_ptr(grape_b200_handle* h, int32_t which); /* 0: grad_J_Tb (local partial), 1: sums[4], 2: pulsevals, 3: G, 4: J_parts[3], 5: tau[K] complex */
/* cudaStream_t the handle launches on (so callers can time with events on it). */
void* grape_b200_stream(grape_b200_handle* h);
/* number of kernels launched by this handle since creation */
int64_t grape_b200_launch_count(const grape_b200_handle* h);
/* Which backward kernels served the last gradient call (instrumentation; synchronises the stream):
 *   0 = GradGenerator block recursion (1+2L operator applications per Taylor order; the only form
 *       of the small / sub-warp paths' :taylor method, of sub-stepped steps and of gradient_method=:taylor),
 *   1 = Krylov form (dense path: chi chain on K columns + one DMMA contraction per step, csrc/dense_kry.cuh).
 *   2 = Krylov form whose strip chains advance two Taylor terms per grid barrier (H_n and H_n^2 strips, csrc/dense.cuh).
 * Both evaluate the same truncated series of the reference's GradGenerator step
 * (src/optimize.jl:880-896, docs/src/background.md:447-494). Negative: error code. */
int grape_b200_gradient_form(grape_b200_handle* h);
/* 1 if the last gradient call ran the forward sweep and the chi chain of the dense path CONCURRENTLY in one
 * cooperative kernel (csrc/dense.cuh dense_chain<2, NS>: chi_k(T) = c_k tgt_k with a scalar c_k, so tgt_k is
 * propagated backwards while Psi_k goes forwards and c_k enters the contraction afterwards; src/optimize.jl:845-855,
 * 880-896), 0 if the sweeps ran one after the other (state running cost, host chi, sub-stepped steps, :taylor).
 * Instrumentation; synchronises the stream. Negative: error code. */
int grape_b200_dense_concurrent(grape_b200_handle* h);
/* Polynomial degree of every time step of the dense path's Krylov-form schedule in the last call (orders[NT]); returns
 * 1 if the chains summed the economised (Chebyshev-cut) polynomial of exp(-i H dt) -- Hermitian generators, csrc/dense.cuh,
 * the reference's Cheby propagator (docs/src/tutorial.md:308, 432) in the monomial basis --, 0 for the Taylor series
 * (the reference's series of src/optimize.jl:604-653). Instrumentation; synchronises. Negative: error code. */
int grape_b200_dense_orders(grape_b200_handle* h, int* orders);
/* Host-only (no GPU): the economised polynomial of degree m (2..20): *theta = the largest ||H dt|| it serves with a
 * uniform error <= 1e-17, g[0..m] = the weights of the Taylor terms (-i H dt)^j / j!.  Test instrumentation. */
int grape_b200_econ_table(int m, double* theta, double* g);
/* Which schedule of the small path (N <= 4) served the last gradient call (instrumentation; synchronises):
 *   0 = not the time-segmented small path (plain chains, sub-warp or dense path),
 *   1 = time-segmented, general generators (forward states read back from fw_storage),
 *   2 = time-segmented, Hermitian generators (forward states recomputed backwards next to chi),
 *   3 = time-segmented, real-symmetric generators (same as 2 in real matrix arithmetic, csrc/small_sym.cuh).
 * All evaluate src/optimize.jl:824-1014 with the same truncation. Negative: error code. */
int grape_b200_small_schedule(grape_b200_handle* h);

/* ---- multi-GPU: peer exchange between shards, one process per GPU -------------------------------------
 * grape_b200_xchg_init allocates this shard's exchange buffer for `world` shards (this one is `rank`) and
 * returns its cudaIpcMemHandle_t (64 bytes) in ipc_handle_out (may be NULL when all shards live in one
 * process).  grape_b200_xchg_attach takes the handles of ALL ranks, [world][64] bytes in rank order, maps the
 * peers' buffers (cudaIpcOpenMemHandle, NVLink/NVSwitch P2P) and switches the evaluation entry points to the
 * sharded form: the k-reduction kernel pushes its partial gradient into every peer's buffer and every rank
 * adds the world partials in rank order (bit-identical on all ranks, no NCCL, no host in the step).
 * The split host API (grape_b200_forward / _backward / _backward_chi) keeps returning LOCAL partials. */
#define GRAPE_B200_IPC_HANDLE_BYTES 64
int grape_b200_xchg_init(grape_b200_handle* h, int32_t rank, int32_t world, void* ipc_handle_out);
int grape_b200_xchg_attach(grape_b200_handle* h, const void* ipc_handles);
int grape_b200_xchg_detach(grape_b200_handle* h);

/* ---- multi-GPU: one host process, several devices (reference: one Julia process, src/optimize.jl:720) --
 * `desc` describes the WHOLE problem (desc->device is ignored); trajectories are split into ndev contiguous
 * blocks, block i on devices[i] (generators unused by a block are not uploaded there).  multi_eval_f /
 * multi_eval_fg = evaluate_functional / evaluate_gradient! of the whole ensemble: one H2D copy of the pulse
 * values per device, all devices run concurrently and meet in the exchange kernels, the result is read from
 * device 0 (tau [2*K] is gathered from all shards). */
typedef struct grape_b200_multi grape_b200_multi;
int  grape_b200_multi_create(const grape_b200_problem* desc, const int32_t* devices, int32_t ndev, grape_b200_multi** out);
void grape_b200_multi_destroy(grape_b200_multi* m);
const char* grape_b200_multi_last_error(const grape_b200_multi* m);   /* m == NULL: last failed multi_create */
int  grape_b200_multi_eval_f(grape_b200_multi* m, const double* pulsevals, double* J_parts, double* tau);
int  grape_b200_multi_eval_fg(grape_b200_multi* m, const double* pulsevals, double* G, double* J_parts,
                              double* tau, double* grad_J_Tb, double* grad_J_a);
int  grape_b200_multi_get_final_states(grape_b200_multi* m, double* out /* [K][N] complex */);
int32_t grape_b200_multi_size(const grape_b200_multi* m);
/* shard i (for the per-shard read-backs above); *k_first = index of its first trajectory */
grape_b200_handle* grape_b200_multi_shard(grape_b200_multi* m, int32_t i, int32_t* k_first);

#ifdef __cplusplus
}
#endif
#endif /* GRAPE_B200_H */
