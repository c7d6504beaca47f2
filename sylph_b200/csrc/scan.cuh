// scan.cuh — small exclusive-scan kernels shared by the post-passes (u32 counters, tens of thousands to a few
// million entries; single-CTA running-carry scan and a two-level scan for the larger arrays).
#pragma once
#include "common.cuh"

namespace syl {

// exclusive scan of n u32 values into out[0..n] (out[n] = total); single CTA, 8 values per thread
// per trip, running carry.  n may come from device memory (d_n != nullptr: n = min(*d_n, n)).
static __global__ void __launch_bounds__(1024) k_scan_u32(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ out,
                                                          const unsigned long long *__restrict__ d_n = nullptr) {
    constexpr int PER = 8;
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (d_n && *d_n < n) n = *d_n;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024 * PER) {
        uint32_t v[PER], tot = 0;
#pragma unroll
        for (int e = 0; e < PER; e++) {
            const uint64_t i = base + (uint64_t)tid * PER + e;
            v[e] = i < n ? in[i] : 0u;
            tot += v[e];
        }
        uint32_t inc = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = wsum[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += t; }
            wsum[lane] = winc - w;  // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        uint32_t run = carry + wsum[wid] + inc - tot;
#pragma unroll
        for (int e = 0; e < PER; e++) {
            const uint64_t i = base + (uint64_t)tid * PER + e;
            if (i < n) out[i] = run;
            run += v[e];
        }
        __syncthreads();
        if (tid == 1023) carry_s = run;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry_s;
}

// two-level exclusive scan for larger arrays: per-CTA local scans + block totals, a single-CTA
// scan of the totals, then the offsets are added back
static __global__ void __launch_bounds__(1024) k_scan_local(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ out,
                                                            uint32_t *__restrict__ block_tot) {
    __shared__ uint32_t wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t i = (uint64_t)blockIdx.x * 1024 + tid;
    const uint32_t v = i < n ? in[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = wsum[lane], winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += t; }
        wsum[lane] = winc - w;
        if (lane == 31) block_tot[blockIdx.x] = winc;
    }
    __syncthreads();
    if (i < n) out[i] = wsum[wid] + inc - v;
}

static __global__ void k_scan_add(uint32_t *__restrict__ out, uint64_t n, const uint32_t *__restrict__ block_off) {
    const uint64_t i = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += block_off[blockIdx.x];
    if (i == n - 1) out[n] = block_off[gridDim.x];  // total
}

// exclusive scan of n u32 values (any n) into out[0..n]: one launch for n <= 64K, three above
static inline int scan_u32(syl_ctx *ctx, const uint32_t *in, uint64_t n, uint32_t *out, DevBuf<uint32_t> &tmp_tot, DevBuf<uint32_t> &tmp_off) {
    cudaStream_t st = ctx->stream;
    if (n <= (1u << 16)) {
        k_scan_u32<<<1, 1024, 0, st>>>(in, n, out);
        ctx->launches++;
        return SYL_OK;
    }
    const uint64_t nb = (n + 1023) / 1024;
    SYL_TRY(tmp_tot.alloc(nb, st));
    SYL_TRY(tmp_off.alloc(nb + 1, st));
    k_scan_local<<<(unsigned)nb, 1024, 0, st>>>(in, n, out, tmp_tot.p);
    k_scan_u32<<<1, 1024, 0, st>>>(tmp_tot.p, nb, tmp_off.p);
    k_scan_add<<<(unsigned)nb, 1024, 0, st>>>(out, n, tmp_off.p);
    ctx->launches += 3;
    return SYL_OK;
}

}  // namespace syl
