// reduce.cuh -- tau / J_T / chi-coefficient and gradient reductions
// (reference src/optimize.jl:752-766, 574-584, 1002-1011). All reductions are
// fixed-order (deterministic run to run).
#pragma once
#include "common.cuh"

GB_D double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

// block-wide sum of NV values per thread; result valid in thread 0. blockDim.x <= 1024.
template <int NV>
GB_D void block_sum(double (&v)[NV], double* s_buf /* [32*NV] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = warp_sum(v[q]);
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NV; ++q) s_buf[warp * NV + q] = v[q];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            double t = lane < nw ? s_buf[lane * NV + q] : 0.0;
            v[q] = warp_sum(t);
        }
    }
    __syncthreads();
}

// sums[0..1] = sum_k w_k tau_k ; sums[2] = sum_k w_k |tau_k|^2 ; sums[3] = sum_k J_b_trajectory[k]
// single block, strided fixed-order accumulation (optimize.jl:752-753, 764-766)
__global__ void __launch_bounds__(1024) reduce_tau(DevP p) {
    __shared__ double s_buf[32 * 4];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
        const double w = p.w ? p.w[k] : 1.0;
        const cplx t = p.tau[k];
        v[0] = fma(w, t.x, v[0]);
        v[1] = fma(w, t.y, v[1]);
        v[2] = fma(w, cnorm2(t), v[2]);
        v[3] += p.jb[k];
    }
    block_sum<4>(v, s_buf);
    if (threadIdx.x == 0) {
        p.sums[0] = v[0]; p.sums[1] = v[1]; p.sums[2] = v[2]; p.sums[3] = v[3];
    }
}

// J_parts from the (global) sums; J_a fluence = sum eps^2 dt.  One block of 256 threads.
// do_tau: the sums of reduce_tau are formed here first (functionals whose chi does not need them before the backward
// sweep: one launch less per gradient).
GB_D void finalize_J_body(const DevP& p, int do_tau, double* s_buf /* [32 * 4] */) {
    if (do_tau) {
        double t4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
            const double w = p.w ? p.w[k] : 1.0;
            const cplx t = p.tau[k];
            t4[0] = fma(w, t.x, t4[0]);
            t4[1] = fma(w, t.y, t4[1]);
            t4[2] = fma(w, cnorm2(t), t4[2]);
            t4[3] += p.jb[k];
        }
        block_sum<4>(t4, s_buf);
        if (threadIdx.x == 0) { p.sums[0] = t4[0]; p.sums[1] = t4[1]; p.sums[2] = t4[2]; p.sums[3] = t4[3]; }
        __syncthreads();
    }
    double v[1] = {0.0};
    if (p.ja_kind == 1) {
        for (int idx = threadIdx.x; idx < p.L * p.NT; idx += blockDim.x) {
            const int n = idx % p.NT;
            const double e = p.eps[idx];
            v[0] = fma(e * e, p.tlist[n + 1] - p.tlist[n], v[0]);
        }
    }
    block_sum<1>(v, s_buf);
    if (threadIdx.x == 0) {
        const double Kg = (double)p.Kglobal;
        double JT;
        if (p.functional == 0) {
            const double fr = p.sums[0] / Kg, fi = p.sums[1] / Kg;
            JT = 1.0 - (fr * fr + fi * fi);
        } else if (p.functional == 1) JT = 1.0 - p.sums[0] / Kg;
        else if (p.functional == 2) JT = 1.0 - p.sums[2] / Kg;
        else JT = nan("");
        p.Jparts[0] = JT;
        p.Jparts[1] = p.ja_kind ? p.lambda_a * v[0] : 0.0;
        p.Jparts[2] = p.gb_kind ? p.lambda_b * p.sums[3] : 0.0;
    }
}
__global__ void __launch_bounds__(256) finalize_J(DevP p, int do_tau) {
    __shared__ double s_buf[32 * 4];
    finalize_J_body(p, do_tau, s_buf);
}

// grad_J_Tb[idx] = -2 * sum_kb partial[kb][idx]   (optimize.jl:574-584)
// grad_J_a[idx]  = 2 eps dt (fluence) ; G = grad_J_Tb + lambda_a grad_J_a  (optimize.jl:1003-1011)
// A block sums 32 gradient elements: warp w takes the slabs kb = w, w + 8, .. (coalesced over idx, the loads of a
// warp are independent), the eight partial sums meet in shared memory in fixed order (deterministic run to run).
// with_J: the last block also does finalize_J's work (one launch less; its J only needs tau / the sums, not the gradient)
__global__ void __launch_bounds__(256) finalize_grad(DevP p, int with_J, int do_tau) {
    __shared__ double s_part[8][32];
    __shared__ double s_bufJ[32 * 4];
    const int LNT = p.L * p.NT;
    const int KB = p.KBdev ? *p.KBdev : p.KB;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = blockIdx.x * 32; base < LNT; base += gridDim.x * 32) {
        const int idx = base + lane;
        double s = 0.0;
        if (idx < LNT) {
#pragma unroll 4
            for (int kb = w; kb < KB; kb += 8) s += p.partial[(size_t)kb * LNT + idx];
        }
        s_part[w][lane] = s;
        __syncthreads();
        if (w == 0 && idx < LNT) {
            double t = s_part[0][lane];
#pragma unroll
            for (int q = 1; q < 8; ++q) t += s_part[q][lane];
            const double gT = -2.0 * t;
            double ga = 0.0;
            if (p.ja_kind == 1) {
                const int n = idx % p.NT;
                ga = 2.0 * p.eps[idx] * (p.tlist[n + 1] - p.tlist[n]);
            }
            p.grad[idx] = p.ja_kind ? fma(p.lambda_a, ga, gT) : gT;
            p.grad[LNT + idx] = gT;
            p.grad[2 * LNT + idx] = ga;
        }
        __syncthreads();
    }
    if (with_J && blockIdx.x == gridDim.x - 1) finalize_J_body(p, do_tau, s_bufJ);
}
