// xchg.cuh -- the two exchanges of a trajectory-sharded gradient evaluation, done by the
// reduction kernels themselves over NVLink peer memory (no NCCL launch, no host in the step).
//
// Trajectories couple only through the functional / chi (all tau_k: reference
// src/optimize.jl:755-760, 845-855) and through the final sum over k
// (_grad_J_T_via_chi!, src/optimize.jl:574-584).  The reference threads over k inside one
// process (`@threadsif`, src/optimize.jl:720, 876); here every shard of trajectories lives
// on its own GPU and owns an EXCHANGE BUFFER that all peers have mapped (cudaIpc* between
// processes, cudaDeviceEnablePeerAccess inside one process):
//
//   slots[ch][parity][rank][XS]   doubles   written by peer `rank` with plain st.global over NVLink
//   flags[ch][rank][XCHG_MAXB]    u64       epoch stamps, one per pushing thread block
//
// Exchange = PUSH: every block stores its chunk of the local partial into the slot
// [parity][my rank] of EVERY peer (posted writes, ~1 us one way), fences at system scope,
// stamps its flag on every peer with the call's epoch, then waits until the same block of
// every peer has stamped the local flags and adds the `world` slots in rank order.  All
// ranks add the same numbers in the same order: the result is bit-identical on every GPU
// and run to run (SURVEY gate G6).  Slots are double buffered on the epoch parity: a peer
// can be at most one exchange ahead (it cannot pass the next wait without our stamp).
//
//   channel 0: the 4 partial sums after the forward sweep (J_T_sm: chi_k needs sum_j tau_j)
//   Epochs are device-side and self-incrementing: an exchange kernel works on (completed epochs + 1) and its last block
//   to finish publishes the new count, so a captured CUDA graph replays and no extra kernel is needed.
//   channel 1: the L*NT partial gradient (+ the 4 sums for J_T_re / J_T_ss, whose chi_k only
//              needs tau_k) fused into finalize_grad
#pragma once
#include "common.cuh"
#include "reduce.cuh"

constexpr int XCHG_MAXW = 16;     // ranks
constexpr int XCHG_MAXB = 296;    // pushing blocks per kernel (2 per SM: all co-resident)

struct XchgDev {
    int rank, world;
    int XS;                               // doubles per slot (>= L*NT + 4, multiple of 2)
    unsigned long long* epoch;            // [2] device-side counters of COMPLETED exchanges (channel 0 / 1): an exchange kernel
                                          // works on epoch + 1 and its last block to finish stores the new value
    unsigned int* ticket;                 // [2] block-completion tickets of the exchange kernels (left at 0)
    double* slots[XCHG_MAXW];             // base of peer r's slot array   [2 ch][2 parity][world][XS]
    unsigned long long* flags[XCHG_MAXW]; // base of peer r's flag array   [2 ch][world][XCHG_MAXB]
    int* timeout;                         // [1] set when a wait gave up (peer never arrived)
};

GB_D size_t xchg_slot_off(const XchgDev& x, int ch, int parity, int rank) {
    return (((size_t)ch * 2 + parity) * x.world + rank) * (size_t)x.XS;
}
GB_D size_t xchg_flag_off(const XchgDev& x, int ch, int rank, int block) {
    return ((size_t)ch * x.world + rank) * XCHG_MAXB + block;
}
GB_D void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
GB_D unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
GB_D double ld_volatile(const double* p) {   // peer-written data: never from a stale L1 line
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
GB_D unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// called by every block when it is done with the exchange: the last one publishes the new epoch for the next launch
GB_D void xchg_finish(const XchgDev& x, int ch, unsigned long long ep) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&x.ticket[ch], 1u) == gridDim.x - 1) {
            x.ticket[ch] = 0;
            x.epoch[ch] = ep;
            __threadfence();
        }
    }
}

// all stores of this block to the peers are done (every thread fenced): stamp + wait
GB_D void xchg_stamp_and_wait(const XchgDev& x, int ch, unsigned long long ep) {
    // The block's peer stores happen-before the barrier; the system-scope RELEASE store of the stamping thread is
    // cumulative over everything ordered before it through the barrier, so one release per peer suffices (a
    // __threadfence_system() by each of the 256 threads cost several NVLink round trips per exchange).
    __syncthreads();
    const int t = threadIdx.x;
    if (t < x.world) {
        st_release_sys(x.flags[t] + xchg_flag_off(x, ch, x.rank, blockIdx.x), ep);   // my stamp on peer t
        const unsigned long long* f = x.flags[x.rank] + xchg_flag_off(x, ch, t, blockIdx.x);
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(f) < ep) {
            if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 4000000000ull) {   // 4 s: a peer died; do not hang the GPU
                *x.timeout = 1;
                break;
            }
        }
    }
    __syncthreads();
}

// channel 0: sums[4] <- sum over ranks, in rank order (single block)
__global__ void __launch_bounds__(32) xchg_sums(DevP p, XchgDev x) {
    const unsigned long long ep = x.epoch[0] + 1;
    const int par = (int)(ep & 1), t = threadIdx.x;
    if (t < 4) {
        const double v = p.sums[t];
        for (int r = 0; r < x.world; ++r) x.slots[r][xchg_slot_off(x, 0, par, x.rank) + t] = v;
    }
    xchg_stamp_and_wait(x, 0, ep);
    if (t < 4) {
        const double* mine = x.slots[x.rank];
        double s = 0.0;
        for (int r = 0; r < x.world; ++r) s += ld_volatile(mine + xchg_slot_off(x, 0, par, r) + t);
        p.sums[t] = s;
    }
    xchg_finish(x, 0, ep);
}

// channel 1, fused with finalize_grad (reduce.cuh): grad_J_Tb = -2 sum_ranks sum_kb partial[kb],
// grad_J_a = 2 eps dt, G = grad_J_Tb + lambda_a grad_J_a (optimize.jl:574-584, 1002-1011);
// with_sums: block 0 also exchanges sums[4] (functionals whose chi does not couple the trajectories); do_tau: it forms
// the local sums from tau first (no reduce_tau launch).  Block 0 finishes with J_parts from the global sums
// (finalize_J's work).  gridDim.x <= XCHG_MAXB; a block owns the 32-element chunks b, b + grid, ...
__global__ void __launch_bounds__(256) finalize_grad_xchg(DevP p, XchgDev x, int with_sums, int do_tau) {
    __shared__ double s_part[8][32];
    __shared__ double s_bufJ[32 * 4];
    const unsigned long long ep = x.epoch[1] + 1;
    const int par = (int)(ep & 1);
    const int LNT = p.L * p.NT;
    const int KB = p.KBdev ? *p.KBdev : p.KB;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t my_slot = xchg_slot_off(x, 1, par, x.rank);
    for (int base = blockIdx.x * 32; base < LNT; base += gridDim.x * 32) {
        const int idx = base + lane;
        double s = 0.0;
        if (idx < LNT) {
#pragma unroll 4
            for (int kb = w; kb < KB; kb += 8) s += p.partial[(size_t)kb * LNT + idx];
        }
        s_part[w][lane] = s;
        __syncthreads();
        if (idx < LNT) {
            double t = s_part[0][lane];
#pragma unroll
            for (int q = 1; q < 8; ++q) t += s_part[q][lane];
            const double gT = -2.0 * t;
            // warp w pushes to the peers w, w + 8, ..: 256-byte coalesced posted writes over NVLink
            for (int r = w; r < x.world; r += 8) x.slots[r][my_slot + idx] = gT;
        }
        __syncthreads();
    }
    if (with_sums && blockIdx.x == 0) {
        if (do_tau) {   // local sums of reduce_tau (same strided fixed-order accumulation)
            double t4[4] = {0.0, 0.0, 0.0, 0.0};
            for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
                const double w = p.w ? p.w[k] : 1.0;
                const cplx t = p.tau[k];
                t4[0] = fma(w, t.x, t4[0]);
                t4[1] = fma(w, t.y, t4[1]);
                t4[2] = fma(w, cnorm2(t), t4[2]);
                t4[3] += p.jb[k];
            }
            block_sum<4>(t4, s_bufJ);
            if (threadIdx.x == 0) { p.sums[0] = t4[0]; p.sums[1] = t4[1]; p.sums[2] = t4[2]; p.sums[3] = t4[3]; }
            __syncthreads();
        }
        if (threadIdx.x < 4) {
            const double v = p.sums[threadIdx.x];
            for (int r = 0; r < x.world; ++r) x.slots[r][my_slot + LNT + threadIdx.x] = v;
        }
    }
    xchg_stamp_and_wait(x, 1, ep);
    const double* mine = x.slots[x.rank];
    for (int base = blockIdx.x * 32; base < LNT; base += gridDim.x * 32) {
        const int idx = base + (threadIdx.x & 31);
        if (w == 0 && idx < LNT) {
            double gT = 0.0;
            for (int r = 0; r < x.world; ++r) gT += ld_volatile(mine + xchg_slot_off(x, 1, par, r) + idx);
            double ga = 0.0;
            if (p.ja_kind == 1) {
                const int n = idx % p.NT;
                ga = 2.0 * p.eps[idx] * (p.tlist[n + 1] - p.tlist[n]);
            }
            p.grad[idx] = p.ja_kind ? fma(p.lambda_a, ga, gT) : gT;
            p.grad[LNT + idx] = gT;
            p.grad[2 * LNT + idx] = ga;
        }
    }
    if (blockIdx.x == 0) {
        if (with_sums && threadIdx.x < 4) {
            double s = 0.0;
            for (int r = 0; r < x.world; ++r) s += ld_volatile(mine + xchg_slot_off(x, 1, par, r) + LNT + threadIdx.x);
            p.sums[threadIdx.x] = s;
        }
        __syncthreads();
        finalize_J_body(p, 0, s_bufJ);   // J_parts from the global sums (finalize_J's work, no extra launch)
    }
    xchg_finish(x, 1, ep);
}
