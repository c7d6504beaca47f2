// sample.cu — sample (read) sketch: seeding survivors -> FxHashMap<Kmer,u32> equivalent.
//
// Replaces the per-record loop of sketch_sequences_needle (src/sketch.rs:917-947):
//   pair_kmer_single (:624-656), extract_markers (:928), and for every survivor
//   dup_removal_lsh_full_exact(.., Some(MAX_DEDUP_COUNT)) (:690-731).
//
// The reference walks reads in file order and keeps ONE global exact set S of (kmer, pair-key)
// plus the count map.  S is keyed by the k-mer, so the state of different k-mers never
// interacts; only the order of the events OF ONE k-mer matters, and that order is read order.
// Device formulation: every survivor becomes an event (hash, read index, p0, p1); events are
// sorted by (hash, read index) with two stable LSD radix sorts; one thread then replays the
// state machine of one k-mer over its (contiguous) event segment.  Within one read the order of
// repeated k-mers is irrelevant (identical events; the second is a duplicate either way).
#include <cub/cub.cuh>

#include <chrono>
#include <cstdlib>
#include <new>
#include <vector>

#include "common.cuh"

namespace syl {

int seed_device(syl_ctx *ctx, const uint8_t *d_bases, uint64_t n_bases, const uint64_t *d_rec_off, uint64_t off_bias,
                uint64_t n_rec, int k, uint64_t c, int sem, int with_pos, syl_survivor *d_out,
                uint64_t cap, uint64_t *n_out);

}  // namespace syl

namespace syl {

constexpr uint64_t NO_PAIR = 1ull;  // bit 0 of recflag set => the read has no pair key

// pair_kmer_single (src/sketch.rs:624-656): four 16-base keys sampled at even/odd offsets from
// the read start and from the middle.  len > 400 (src/sketch.rs:923) or len < 66 (:627) => None.
// One thread per survivor: the 32 bytes a key pair is drawn from are fetched as nine aligned
// 32-bit words and realigned with funnel shifts; even / odd bytes are separated with PRMT and
// mapped through four pre-shifted copies of the exact BYTE_TO_SEQ table in shared memory.
constexpr int EV_THREADS = 128;

__device__ __forceinline__ void load32_unaligned(const uint8_t *p, uint32_t x[8]) {
    const uintptr_t ad = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(ad & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(ad & 3u) * 8u;
    uint32_t v[9];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __ldg(w + i);
    v[8] = sh ? __ldg(w + 8) : 0u;  // an aligned word holding a valid byte never leaves the buffer's last word
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __funnelshift_r(v[i], v[i + 1], sh);
}

// 16 bytes picked by `sel` (0x6420 = even, 0x7531 = odd) out of 32 -> 16 two-bit codes, MSB-first
__device__ __forceinline__ uint32_t pack16(const uint32_t x[8], uint32_t sel, const uint8_t (*lut)[256]) {
    uint32_t out = 0;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const uint32_t e = __byte_perm(x[2 * m], x[2 * m + 1], sel);
        const uint32_t g = ((uint32_t)lut[0][e & 0xFFu] | (uint32_t)lut[1][(e >> 8) & 0xFFu] |
                            (uint32_t)lut[2][(e >> 16) & 0xFFu]) | (uint32_t)lut[3][e >> 24];
        out = (out << 8) | g;
    }
    return out;
}

__global__ void __launch_bounds__(EV_THREADS)
k_events(const syl_survivor *__restrict__ sv, uint64_t n, const uint8_t *__restrict__ bases,
         const uint64_t *__restrict__ rec_off, uint64_t off_bias, uint64_t rec_base, int no_dedup,
         uint64_t *__restrict__ hash, uint64_t *__restrict__ recflag,
         uint64_t *__restrict__ p0, uint64_t *__restrict__ p1) {
    __shared__ uint8_t lut[4][256];
    for (int i = threadIdx.x; i < 256; i += EV_THREADS) {
        const uint32_t code = byte_to_seq((uint32_t)i);
        lut[0][i] = (uint8_t)(code << 6);
        lut[1][i] = (uint8_t)(code << 4);
        lut[2][i] = (uint8_t)(code << 2);
        lut[3][i] = (uint8_t)code;
    }
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * EV_THREADS + threadIdx.x;
    if (i >= n) return;
    const syl_survivor s = sv[i];
    hash[i] = s.hash;
    const uint64_t a = rec_off[s.rec] - off_bias;
    const uint64_t L = rec_off[s.rec + 1] - off_bias - a;
    const bool has_pair = !no_dedup && L <= 400 && L >= 66;
    recflag[i] = ((rec_base + s.rec) << 1) | (has_pair ? 0ull : NO_PAIR);
    uint64_t k0 = 0, k1 = 0;
    if (has_pair) {
        uint32_t x[8];
        load32_unaligned(bases + a, x);
        const uint32_t f = pack16(x, 0x6420, lut), g = pack16(x, 0x7531, lut);
        load32_unaligned(bases + a + L / 2, x);
        const uint32_t r = pack16(x, 0x6420, lut), t = pack16(x, 0x7531, lut);
        k0 = ((uint64_t)f << 32) | r;  // doublepairs.0 = [kmer_f, kmer_r]
        k1 = ((uint64_t)g << 32) | t;  // doublepairs.1 = [kmer_g, kmer_t]
    }
    p0[i] = k0;
    p1[i] = k1;
}

__global__ void k_iota(uint32_t *idx, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}

template <typename T>
__global__ void k_gather(const T *__restrict__ src, const uint32_t *__restrict__ idx, T *__restrict__ dst, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// One thread replays dup_removal_lsh_full_exact (src/sketch.rs:690-731) for one k-mer.
// set[] is this segment's private slice of a global scratch array (2 slots per event).
__global__ void k_dedup(const uint64_t *__restrict__ seg_off, const uint32_t *__restrict__ seg_len, uint64_t n_seg,
                        const uint32_t *__restrict__ order, const uint64_t *__restrict__ recflag,
                        const uint64_t *__restrict__ p0, const uint64_t *__restrict__ p1,
                        uint64_t *__restrict__ set, uint32_t *__restrict__ count,
                        unsigned long long *__restrict__ n_dup) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const uint64_t start = seg_off[s];
    const uint32_t len = seg_len[s];
    if (len == 1) {  // first occurrence is never a duplicate (c == 0)
        count[s] = 1;
        return;
    }
    uint64_t *S = set + 2 * start;
    uint32_t nset = 0, c = 0, dups = 0;
    for (uint32_t e = 0; e < len; e++) {
        if (c >= 4u) {  // MAX_DEDUP_COUNT (src/constants.rs:14): dedup is off from here on
            c += len - e;
            break;
        }
        const uint32_t ev = order[start + e];
        if (recflag[ev] & NO_PAIR) {
            c++;
            continue;
        }
        const uint64_t a = p0[ev], b = p1[ev];
        bool ret = false, found = false;
        for (uint32_t q = 0; q < nset; q++) found |= (S[q] == a);
        if (found) ret = c > 0; else S[nset++] = a;
        found = false;
        for (uint32_t q = 0; q < nset; q++) found |= (S[q] == b);
        if (found) ret = ret || c > 0; else S[nset++] = b;
        if (ret) dups++; else c++;
    }
    count[s] = c;
    if (dups) atomicAdd(n_dup, (unsigned long long)dups);
}

__global__ void k_copy_len(const uint32_t *__restrict__ len, uint32_t *__restrict__ count, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) count[i] = len[i];
}

static inline unsigned nblk(uint64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

static inline int bits_for(uint64_t maxval) {
    int b = 1;
    while (b < 64 && (maxval >> b)) b++;
    return b;
}

// Accumulates events over one or more seeding batches, then finishes into a syl_sample.
struct SampleBuilder {
    syl_ctx *ctx;
    int k;
    uint64_t c;
    int no_dedup, sem;
    uint64_t n_reads = 0, n_bases = 0, n_events = 0, cap = 0;
    uint64_t *hash = nullptr, *recflag = nullptr, *p0 = nullptr, *p1 = nullptr;

    ~SampleBuilder() { release(); }
    void release() {
        cudaStream_t st = ctx->stream;
        if (hash) cudaFreeAsync(hash, st);
        if (recflag) cudaFreeAsync(recflag, st);
        if (p0) cudaFreeAsync(p0, st);
        if (p1) cudaFreeAsync(p1, st);
        hash = recflag = p0 = p1 = nullptr;
        cap = 0;
    }
    int reserve(uint64_t need) {
        if (need <= cap) return SYL_OK;
        uint64_t ncap = std::max<uint64_t>(need, cap * 2);
        cudaStream_t st = ctx->stream;
        uint64_t *nh, *nr, *n0, *n1;
        SYL_CUDA(cudaMallocAsync((void **)&nh, ncap * 8, st));
        SYL_CUDA(cudaMallocAsync((void **)&nr, ncap * 8, st));
        SYL_CUDA(cudaMallocAsync((void **)&n0, ncap * 8, st));
        SYL_CUDA(cudaMallocAsync((void **)&n1, ncap * 8, st));
        if (n_events) {
            SYL_CUDA(cudaMemcpyAsync(nh, hash, n_events * 8, cudaMemcpyDeviceToDevice, st));
            SYL_CUDA(cudaMemcpyAsync(nr, recflag, n_events * 8, cudaMemcpyDeviceToDevice, st));
            SYL_CUDA(cudaMemcpyAsync(n0, p0, n_events * 8, cudaMemcpyDeviceToDevice, st));
            SYL_CUDA(cudaMemcpyAsync(n1, p1, n_events * 8, cudaMemcpyDeviceToDevice, st));
        }
        release();
        hash = nh; recflag = nr; p0 = n0; p1 = n1;
        cap = ncap;
        return SYL_OK;
    }

    // one batch of reads, device resident; read indices continue from the previous batch
    int add(const uint8_t *d_bases, uint64_t nb, const uint64_t *d_off, uint64_t off_bias, uint64_t nr) {
        cudaStream_t st = ctx->stream;
        if (nr == 0) return SYL_OK;
        uint64_t scap = nb / c + nb / (4 * c) + 65536;
        if (scap > nb) scap = nb + 16;
        DevBuf<syl_survivor> sv;
        uint64_t n = 0;
        for (;;) {
            SYL_TRY(sv.alloc(scap, st));
            int rc = seed_device(ctx, d_bases, nb, d_off, off_bias, nr, k, c, sem, /*with_pos=*/0, sv.p, scap, &n);
            if (rc == SYL_ERR_CAPACITY) { scap = n + 16; continue; }
            if (rc != SYL_OK) return rc;
            break;
        }
        SYL_TRY(reserve(n_events + n));
        if (n) {
            k_events<<<nblk(n, EV_THREADS), EV_THREADS, 0, st>>>(sv.p, n, d_bases, d_off, off_bias, n_reads, no_dedup, hash + n_events,
                                                    recflag + n_events, p0 + n_events, p1 + n_events);
            ctx->launches++;
            SYL_CUDA(cudaGetLastError());
        }
        n_events += n;
        n_reads += nr;
        n_bases += nb;
        return SYL_OK;
    }

    int finish(syl_sample **out) {
        cudaStream_t st = ctx->stream;
        syl_sample *s = new (std::nothrow) syl_sample();
        if (!s) return SYL_ERR_OOM;
        s->device = ctx->device;
        s->stream = ctx->stream;
        s->k = k;
        s->c = c;
        s->mean_read_length = n_reads ? (double)n_bases / (double)n_reads : 0.;
        const uint64_t N = n_events;
        if (N == 0) { *out = s; return SYL_OK; }
        if (N >= 0xFFFFFFFFull) { delete s; set_error("more than 2^32-2 survivor events in one sample"); return SYL_ERR_ARG; }

        // -- order events by (hash, read): sort by read first, then stable sort by hash
        DevBuf<uint32_t> idx_a, idx_b;
        DevBuf<uint64_t> key_a, key_b;
        SYL_TRY(idx_a.alloc(N, st)); SYL_TRY(idx_b.alloc(N, st));
        SYL_TRY(key_a.alloc(N, st)); SYL_TRY(key_b.alloc(N, st));
        k_iota<<<nblk(N, 256), 256, 0, st>>>(idx_a.p, N);
        ctx->launches++;
        const uint64_t thr = fmh_threshold(c);
        const int hash_bits = bits_for(thr);
        DevBuf<uint8_t> tmp;
        size_t tmp_bytes = 0, t2 = 0;
        uint32_t *ord = idx_a.p;  // final event order
        uint64_t *hs = key_a.p;   // hashes in final order
        if (!no_dedup) {
            const int rec_bits = bits_for((n_reads << 1) | 1);
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, recflag, key_b.p, idx_a.p, idx_b.p, N, 0, rec_bits, st);
            cub::DeviceRadixSort::SortPairs(nullptr, t2, key_a.p, key_b.p, idx_b.p, idx_a.p, N, 0, hash_bits, st);
            tmp_bytes = std::max(tmp_bytes, t2);
            SYL_TRY(tmp.alloc(tmp_bytes, st));
            SYL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, recflag, key_b.p, idx_a.p, idx_b.p, N, 0, rec_bits, st));
            k_gather<uint64_t><<<nblk(N, 256), 256, 0, st>>>(hash, idx_b.p, key_a.p, N);
            SYL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_a.p, key_b.p, idx_b.p, idx_a.p, N, 0, hash_bits, st));
            ctx->launches += 2 + 2 * 8;
            hs = key_b.p;
            ord = idx_a.p;
        } else {
            cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, hash, key_b.p, N, 0, hash_bits, st);
            SYL_TRY(tmp.alloc(tmp_bytes, st));
            SYL_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, hash, key_b.p, N, 0, hash_bits, st));
            ctx->launches += 8;
            hs = key_b.p;
        }

        // -- segments = distinct hashes
        DevBuf<uint64_t> uniq, seg_off;
        DevBuf<uint32_t> seg_len;
        SYL_TRY(uniq.alloc(N, st)); SYL_TRY(seg_off.alloc(N + 1, st)); SYL_TRY(seg_len.alloc(N, st));
        uint64_t *d_nruns = ctx->d_counters + 1;
        size_t rle_bytes = 0, scan_bytes = 0;
        cub::DeviceRunLengthEncode::Encode(nullptr, rle_bytes, hs, uniq.p, seg_len.p, d_nruns, N, st);
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, seg_len.p, seg_off.p, N, st);
        DevBuf<uint8_t> tmp2;
        SYL_TRY(tmp2.alloc(std::max(rle_bytes, scan_bytes), st));
        size_t tb = std::max(rle_bytes, scan_bytes);
        SYL_CUDA(cub::DeviceRunLengthEncode::Encode(tmp2.p, tb, hs, uniq.p, seg_len.p, d_nruns, N, st));
        SYL_CUDA(cudaMemcpyAsync(ctx->h_counters + 1, d_nruns, 8, cudaMemcpyDeviceToHost, st));
        SYL_CUDA(cudaStreamSynchronize(st));
        const uint64_t U = ctx->h_counters[1];
        ctx->launches += 2;
        SYL_CUDA(cudaMallocAsync((void **)&s->hash, std::max<uint64_t>(U, 1) * 8, st));
        SYL_CUDA(cudaMallocAsync((void **)&s->count, std::max<uint64_t>(U, 1) * 4, st));
        SYL_CUDA(cudaMemcpyAsync(s->hash, uniq.p, U * 8, cudaMemcpyDeviceToDevice, st));
        s->n = U;
        if (no_dedup) {
            k_copy_len<<<nblk(U, 256), 256, 0, st>>>(seg_len.p, s->count, U);
            ctx->launches++;
        } else {
            tb = std::max(rle_bytes, scan_bytes);
            SYL_CUDA(cub::DeviceScan::ExclusiveSum(tmp2.p, tb, seg_len.p, seg_off.p, U, st));
            DevBuf<uint64_t> set;
            SYL_TRY(set.alloc(2 * N, st));
            unsigned long long *d_ndup = reinterpret_cast<unsigned long long *>(ctx->d_counters + 2);
            SYL_CUDA(cudaMemsetAsync(d_ndup, 0, 8, st));
            k_dedup<<<nblk(U, 128), 128, 0, st>>>(seg_off.p, seg_len.p, U, ord, recflag, p0, p1, set.p, s->count, d_ndup);
            ctx->launches += 2;
            SYL_CUDA(cudaGetLastError());
            SYL_CUDA(cudaMemcpyAsync(ctx->h_counters + 2, d_ndup, 8, cudaMemcpyDeviceToHost, st));
            SYL_CUDA(cudaStreamSynchronize(st));
            s->num_dup_removed = ctx->h_counters[2];
        }
        SYL_CUDA(cudaStreamSynchronize(st));
        *out = s;
        return SYL_OK;
    }
};

}  // namespace syl

using namespace syl;

extern "C" {

int syl_sketch_reads(syl_ctx *ctx, int mem, const uint8_t *bases, uint64_t n_bases,
                     const uint64_t *rec_off, uint64_t n_reads, int k, uint64_t c, int no_dedup,
                     int sem, syl_sample **out) {
    if (!ctx || !out || (!bases && n_bases) || !rec_off) { set_error("NULL argument"); return SYL_ERR_ARG; }
    if (c == 0) { set_error("c must be >= 1"); return SYL_ERR_ARG; }
    *out = nullptr;
    SYL_CUDA(cudaSetDevice(ctx->device));
    SampleBuilder b{ctx, k, c, no_dedup, sem};
    cudaStream_t st = ctx->stream;
    if (mem == SYL_MEM_DEVICE) {
        SYL_TRY(b.add(bases, n_bases, rec_off, 0, n_reads));
        return b.finish(out);
    }
    if (mem != SYL_MEM_HOST) { set_error("bad mem"); return SYL_ERR_ARG; }
    // Host buffers: cut the reads into chunks of <= CHUNK bytes on record boundaries and copy
    // chunk i+1 on the ctx copy stream while chunk i is being seeded (pinned caller memory
    // overlaps fully).  Bases and the matching slice of rec_off are copied verbatim; kernels
    // subtract the slice's first offset (off_bias).  Staging buffers live in the ctx.
    const uint64_t CHUNK = 128ull << 20;
    static const bool dbg = getenv("SYL_DEBUG_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    if (!ctx->copy_stream) {
        SYL_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            SYL_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
            SYL_CUDA(cudaEventCreateWithFlags(&ctx->ev_used[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t cs = ctx->copy_stream;
    int rc = SYL_OK;
    uint64_t r0 = 0;
    int slot = 0;
    struct Pending { uint64_t nb, nr, bias; int slot; bool valid; } pend = {0, 0, 0, 0, false};
    while (r0 < n_reads || pend.valid) {
        Pending next = {0, 0, 0, slot, false};
        if (r0 < n_reads) {
            // records [r0, r1) with total bytes <= CHUNK (at least one record)
            uint64_t lo = r0 + 1, hi = n_reads;
            const uint64_t base = rec_off[r0];
            while (lo < hi) {
                uint64_t mid = (lo + hi + 1) >> 1;
                if (rec_off[mid] - base <= CHUNK) lo = mid; else hi = mid - 1;
            }
            const uint64_t r1 = lo, nb = rec_off[r1] - base, nr = r1 - r0;
            if (nb + 64 > ctx->stage_cap_b[slot]) {
                cudaStreamSynchronize(st);
                cudaStreamSynchronize(cs);
                if (ctx->stage_b[slot]) cudaFree(ctx->stage_b[slot]);
                ctx->stage_b[slot] = nullptr;
                ctx->stage_cap_b[slot] = std::max<uint64_t>(nb + 64, CHUNK + 64);
                if (cudaMalloc((void **)&ctx->stage_b[slot], ctx->stage_cap_b[slot]) != cudaSuccess) {
                    ctx->stage_cap_b[slot] = 0; rc = SYL_ERR_OOM; set_error("staging alloc failed"); break;
                }
            }
            if (nr + 1 > ctx->stage_cap_o[slot]) {
                cudaStreamSynchronize(st);
                cudaStreamSynchronize(cs);
                if (ctx->stage_o[slot]) cudaFree(ctx->stage_o[slot]);
                ctx->stage_o[slot] = nullptr;
                ctx->stage_cap_o[slot] = (nr + 1) * 2;
                if (cudaMalloc((void **)&ctx->stage_o[slot], ctx->stage_cap_o[slot] * 8) != cudaSuccess) {
                    ctx->stage_cap_o[slot] = 0; rc = SYL_ERR_OOM; set_error("staging alloc failed"); break;
                }
            }
            cudaStreamWaitEvent(cs, ctx->ev_used[slot], 0);  // previous user of this slot is done
            if (cudaMemcpyAsync(ctx->stage_b[slot], bases + base, nb, cudaMemcpyHostToDevice, cs) != cudaSuccess ||
                cudaMemcpyAsync(ctx->stage_o[slot], rec_off + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, cs) != cudaSuccess) {
                rc = SYL_ERR_CUDA; set_error("H2D copy failed"); break;
            }
            cudaEventRecord(ctx->ev_copied[slot], cs);
            next = {nb, nr, base, slot, true};
            r0 = r1;
            slot ^= 1;
        }
        if (pend.valid) {
            cudaStreamWaitEvent(st, ctx->ev_copied[pend.slot], 0);
            rc = b.add(ctx->stage_b[pend.slot], pend.nb, ctx->stage_o[pend.slot], pend.bias, pend.nr);
            cudaEventRecord(ctx->ev_used[pend.slot], st);
            if (rc != SYL_OK) break;
        }
        pend = next;
    }
    const double t_loop = now() - t_start, t0 = now();
    if (rc == SYL_OK) rc = b.finish(out);
    else { cudaStreamSynchronize(cs); cudaStreamSynchronize(st); }
    if (dbg) fprintf(stderr, "[syl_sketch_reads host] chunk loop %.2f ms, finish %.2f ms\n", t_loop, now() - t0);
    return rc;
}

int syl_sample_upload(syl_ctx *ctx, int mem, const uint64_t *hash, const uint32_t *count,
                      uint64_t n, int k, uint64_t c, syl_sample **out) {
    if (!ctx || !out || (n && (!hash || !count))) { set_error("NULL argument"); return SYL_ERR_ARG; }
    *out = nullptr;
    SYL_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    syl_sample *s = new (std::nothrow) syl_sample();
    if (!s) return SYL_ERR_OOM;
    s->device = ctx->device; s->stream = st; s->k = k; s->c = c; s->n = n;
    SYL_CUDA(cudaMallocAsync((void **)&s->hash, std::max<uint64_t>(n, 1) * 8, st));
    SYL_CUDA(cudaMallocAsync((void **)&s->count, std::max<uint64_t>(n, 1) * 4, st));
    if (n) {
        DevBuf<uint64_t> kin;
        DevBuf<uint32_t> vin;
        SYL_TRY(kin.alloc(n, st)); SYL_TRY(vin.alloc(n, st));
        cudaMemcpyKind kind = mem == SYL_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        SYL_CUDA(cudaMemcpyAsync(kin.p, hash, n * 8, kind, st));
        SYL_CUDA(cudaMemcpyAsync(vin.p, count, n * 4, kind, st));
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, kin.p, s->hash, vin.p, s->count, n, 0, 64, st);
        DevBuf<uint8_t> tmp;
        SYL_TRY(tmp.alloc(tb, st));
        SYL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, kin.p, s->hash, vin.p, s->count, n, 0, 64, st));
        ctx->launches += 8;
        SYL_CUDA(cudaStreamSynchronize(st));
    }
    *out = s;
    return SYL_OK;
}

uint64_t syl_sample_size(const syl_sample *s) { return s ? s->n : 0; }
double syl_sample_mean_read_length(const syl_sample *s) { return s ? s->mean_read_length : 0.; }
uint64_t syl_sample_num_dup_removed(const syl_sample *s) { return s ? s->num_dup_removed : 0; }

int syl_sample_download(syl_ctx *ctx, const syl_sample *s, uint64_t *hash, uint32_t *count) {
    if (!ctx || !s) { set_error("NULL argument"); return SYL_ERR_ARG; }
    SYL_CUDA(cudaSetDevice(ctx->device));
    if (s->n && hash) SYL_CUDA(cudaMemcpyAsync(hash, s->hash, s->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (s->n && count) SYL_CUDA(cudaMemcpyAsync(count, s->count, s->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SYL_CUDA(cudaStreamSynchronize(ctx->stream));
    return SYL_OK;
}

int syl_sample_device_ptrs(const syl_sample *s, const uint64_t **hash, const uint32_t **count) {
    if (!s) { set_error("NULL argument"); return SYL_ERR_ARG; }
    if (hash) *hash = s->hash;
    if (count) *count = s->count;
    return SYL_OK;
}

void syl_sample_free(syl_sample *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    // stream-ordered free on the creating ctx stream (free samples before destroying their ctx)
    if (s->hash) cudaFreeAsync(s->hash, s->stream);
    if (s->count) cudaFreeAsync(s->count, s->stream);
    delete s;
}

}  // extern "C"
