// sample.cu — sample (read) sketch: seeding survivors -> FxHashMap<Kmer,u32> equivalent.
//
// Replaces the per-record loop of sketch_sequences_needle (src/sketch.rs:917-947):
//   pair_kmer_single (:624-656), extract_markers (:928), and for every survivor
//   dup_removal_lsh_full_exact(.., Some(MAX_DEDUP_COUNT)) (:690-731).
//
// The reference walks reads in file order and keeps ONE global exact set S of (kmer, pair-key)
// plus the count map.  S is keyed by the k-mer, so the state of different k-mers never
// interacts; only the order of the events OF ONE k-mer matters, and that order is read order.
// Device formulation: every survivor is an event (hash, read index, p0, p1), written by the
// seeding kernel itself; the events are partitioned by hash bucket, a CTA groups the events of
// ~25 consecutive buckets by k-mer in shared memory and one thread replays the state machine
// of one k-mer over its first events in read order (after four counted events every further
// event counts).  Groups that do not fit take the generic path: two stable LSD radix sorts by
// (hash, read index) and one thread per (contiguous) k-mer segment.  Within one read the order
// of repeated k-mers is irrelevant (identical events; the second is a duplicate either way).
#include <cub/cub.cuh>

#include <chrono>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace syl {

int seed_device_ex(syl_ctx *ctx, const uint8_t *d_bases, uint64_t n_bases, const uint64_t *d_rec_off, uint64_t off_bias,
                   uint64_t n_rec, int k, uint64_t c, int sem, int with_pos, void *d_out,
                   uint64_t cap, uint64_t *n_out, int emit_events, uint64_t rec_base, int no_dedup,
                   uint32_t *d_pend, uint64_t *n_pend, uint32_t *d_bucket_cnt, uint64_t Mb, uint32_t nbk);

}  // namespace syl

namespace syl {

// pair_kmer_single (src/sketch.rs:624-656): four 16-base keys sampled at even/odd offsets from
// the read start and from the middle.  len > 400 (src/sketch.rs:923) or len < 66 (:627) => None.
// k_seed computes the keys itself from its packed tile whenever the read's first 32 bases and the
// 32 bases from its middle are inside the tile; k_events_fix below handles the reads cut by a tile
// edge (about 1 % of 150 bp reads).  One thread per such event: the 32 bytes a key pair is drawn
// from are fetched as nine aligned 32-bit words and realigned with funnel shifts; even / odd bytes
// are separated with PRMT and mapped through four pre-shifted copies of the exact BYTE_TO_SEQ
// table in shared memory.
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
k_events_fix(EventRec *__restrict__ ev, const uint32_t *__restrict__ pend, const unsigned long long *__restrict__ n_pend,
             const uint8_t *__restrict__ bases, const uint64_t *__restrict__ rec_off, uint64_t off_bias, uint64_t rec_base) {
    __shared__ uint8_t lut[4][256];
    for (int i = threadIdx.x; i < 256; i += EV_THREADS) {
        const uint32_t code = byte_to_seq((uint32_t)i);
        lut[0][i] = (uint8_t)(code << 6);
        lut[1][i] = (uint8_t)(code << 4);
        lut[2][i] = (uint8_t)(code << 2);
        lut[3][i] = (uint8_t)code;
    }
    __syncthreads();
    const uint64_t n = *n_pend;
    for (uint64_t i = (uint64_t)blockIdx.x * EV_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * EV_THREADS) {
        EventRec *e = ev + pend[i];
        const uint64_t rf = e->recflag & ~EV_PENDING;
        const uint64_t rec = (rf >> 1) - rec_base;
        const uint64_t a = rec_off[rec] - off_bias;
        const uint64_t L = rec_off[rec + 1] - off_bias - a;
        uint32_t x[8];
        load32_unaligned(bases + a, x);
        const uint32_t f = pack16(x, 0x6420, lut), g = pack16(x, 0x7531, lut);
        load32_unaligned(bases + a + L / 2, x);
        const uint32_t r = pack16(x, 0x6420, lut), t = pack16(x, 0x7531, lut);
        e->recflag = rf;
        e->p0 = ((uint64_t)f << 32) | r;  // doublepairs.0 = [kmer_f, kmer_r]
        e->p1 = ((uint64_t)g << 32) | t;  // doublepairs.1 = [kmer_g, kmer_t]
    }
}

__global__ void k_unpack_events(const EventRec *__restrict__ ev, uint64_t n, uint64_t *__restrict__ hash,
                                uint64_t *__restrict__ recflag, uint64_t *__restrict__ p0, uint64_t *__restrict__ p1) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const EventRec r = ev[i];
    hash[i] = r.hash; recflag[i] = r.recflag; p0[i] = r.p0; p1[i] = r.p1;
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


// ------------------------------------------------------------------------------------------------
// Post-pass, primary path: MSD bucket partition + CTA-local sort + fused run-length / dedup.
//   k_bucket_hist / k_scan_u32 / k_scatter_events : partition the events by bucket =
//       mulhi(hash, Mb) (monotone in the hash, uniform because hashes are uniform below thr)
//   k_group_dedup : one CTA takes the consecutive buckets whose start offset falls into
//       [g*T, (g+1)*T) (<= GRP_CAP events), bitonic-sorts them by (hash, read) in shared memory,
//       finds the k-mer segments and replays dup_removal_lsh_full_exact per segment
//   k_compact_uniq : staged (hash,count) pairs -> dense arrays; buckets are monotone in the hash
//       and every group is sorted, so the result is globally sorted without a merge.
// Groups that do not fit (heavy-hitter k-mers, > GRP_CAP events) or whose dedup set outgrows
// the per-thread buffer are handed to the generic radix-sort path below and merged at the end.
constexpr int GRP_THREADS = 256;
constexpr int GRP_CAP = 1024;   // most events one CTA handles in shared memory
constexpr int GRP_T = 768;      // group span in event offsets: typical n ~ 800, leaving room for k-mers with ~200 events

__global__ void k_bucket_hist(const EventRec *__restrict__ ev, uint64_t n, uint64_t Mb, uint32_t nbk,
                              uint32_t *__restrict__ cnt) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t b = (uint32_t)__umul64hi(ev[i].hash, Mb);
    atomicAdd(&cnt[b < nbk ? b : nbk - 1], 1u);
}

// exclusive scan of n u32 values into out[0..n] (out[n] = total); single CTA, 8 values per thread
// per trip, running carry
__global__ void __launch_bounds__(1024) k_scan_u32(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ out) {
    constexpr int PER = 8;
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
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
__global__ void __launch_bounds__(1024) k_scan_local(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ out,
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

__global__ void k_scan_add(uint32_t *__restrict__ out, uint64_t n, const uint32_t *__restrict__ block_off) {
    const uint64_t i = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += block_off[blockIdx.x];
    if (i == n - 1) out[n] = block_off[gridDim.x];  // total
}

__global__ void k_scatter_events(const EventRec *__restrict__ ev, uint64_t n,
                                 uint64_t Mb, uint32_t nbk, const uint32_t *__restrict__ boff,
                                 uint32_t *__restrict__ cursor, EventRec *__restrict__ part) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(ev + i);
    const uint4 lo = __ldcs(src), hi = __ldcs(src + 1);  // read once: keep L2 for the scattered writes
    const uint64_t h = ((uint64_t)lo.y << 32) | lo.x;
    uint32_t b = (uint32_t)__umul64hi(h, Mb);
    if (b >= nbk) b = nbk - 1;
    const uint32_t pos = boff[b] + atomicAdd(&cursor[b], 1u);
    uint4 *dst = reinterpret_cast<uint4 *>(part + pos);
    dst[0] = lo;
    dst[1] = hi;
}

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *a, uint32_t n, uint64_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t)a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// group g = the consecutive buckets whose start offset lies in [g*T, (g+1)*T): bucket range -> g_bf / g_be
__global__ void k_group_ranges(const uint32_t *__restrict__ boff, uint32_t nbk, uint32_t ng, uint32_t *__restrict__ g_bf,
                               uint32_t *__restrict__ g_be) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ng) return;
    uint32_t bf = lower_bound_u32(boff, nbk + 1, (uint64_t)g * GRP_T);
    uint32_t be = lower_bound_u32(boff, nbk + 1, (uint64_t)(g + 1) * GRP_T);
    if (bf > nbk) bf = nbk;
    if (be > nbk) be = nbk;
    if (be < bf) be = bf;
    g_bf[g] = bf;
    g_be[g] = be;
}

__device__ __forceinline__ uint32_t bucket_of(uint64_t h, uint64_t Mb, uint32_t nbk) {
    const uint32_t b = (uint32_t)__umul64hi(h, Mb);
    return b < nbk ? b : nbk - 1;
}

// One CTA per group (<= GRP_CAP events, consecutive buckets). Shared-memory traffic is what
// bounds this kernel, so instead of sorting all events (a 128-bit-key bitonic sort was 4x slower)
// it (1) groups equal hashes with an open-addressing table, (2) replays
// dup_removal_lsh_full_exact on each k-mer's first events in read order, (3) orders the resulting
// unique (hash, count) pairs: the group's buckets are already monotone in the hash, so a pair's
// position is its bucket's offset plus its rank among the ~10 pairs of the same bucket (a bitonic
// sort of the pairs remains for groups spanning more than GRP_LB buckets).
constexpr int GRP_SLOTS = 2048;  // table slots (load factor <= 0.5)
constexpr int GRP_LB = 512;      // most buckets per group for the bucket-rank output ordering (else bitonic sort)
constexpr int GRP_SELECT_STEPS = 64;  // selection steps before a duplicate-heavy k-mer goes to the generic path

struct GroupSmem {
    unsigned long long ht[GRP_SLOTS];  // 16 KB  hash per slot
    uint64_t rf[GRP_CAP];              //  8 KB  recflag per event        | after the replay: unique hashes
    uint64_t p0[GRP_CAP];              //  8 KB  first pair key per event | after the replay: unique counts (u32)
    uint64_t p1[GRP_CAP];              //  8 KB  second pair key per event
    uint16_t scnt[GRP_SLOTS];          //  4 KB  events per slot, then fill cursor, then the k-mer's count
    uint16_t soff[GRP_SLOTS + 2];      //  4 KB  exclusive scan of scnt
    uint16_t ev_slot[GRP_CAP];         //  2 KB  slot per event | after the member fill: 2 x GRP_LB bucket counters / offsets
    uint16_t member[GRP_CAP];          //  2 KB  event indices grouped by slot
    uint16_t occ[GRP_CAP];             //  2 KB  occupied slots, compacted
    uint32_t wtot[GRP_THREADS / 32], wocc[GRP_THREADS / 32];
    uint32_t e0, e1, overflow, dups, nuniq, nlong, lb_n, bf;
    uint16_t longs[GRP_CAP / 2];      //  1 KB  slots of k-mers that need the warp-cooperative replay
};

__global__ void __launch_bounds__(GRP_THREADS)
k_group_dedup(const EventRec *__restrict__ part, const uint32_t *__restrict__ boff, const uint32_t *__restrict__ g_bf,
              const uint32_t *__restrict__ g_be, uint32_t cap, uint64_t Mb, uint32_t nbk,
              int no_dedup, uint64_t *__restrict__ st_hash, uint32_t *__restrict__ st_cnt,
              uint32_t *__restrict__ g_nuniq, uint32_t *__restrict__ g_e0, uint32_t *__restrict__ g_n,
              uint8_t *__restrict__ g_fallback, unsigned long long *__restrict__ n_dup) {
    extern __shared__ __align__(16) uint8_t grp_smem_raw[];
    GroupSmem &S = *reinterpret_cast<GroupSmem *>(grp_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t g = blockIdx.x;
    if (tid == 0) {
        S.e0 = boff[g_bf[g]];  // boff[nbk] = N
        S.e1 = boff[g_be[g]];
        S.bf = g_bf[g];
        S.lb_n = g_be[g] - g_bf[g];
        S.overflow = 0;
        S.dups = 0;
        S.nlong = 0;
    }
    for (int i = tid; i < GRP_SLOTS; i += GRP_THREADS) { S.ht[i] = 0xFFFFFFFFFFFFFFFFull; S.scnt[i] = 0; }
    __syncthreads();
    const uint32_t e0 = S.e0, n = S.e1 - S.e0, bf = S.bf;
    if (tid == 0) { g_e0[g] = e0; g_n[g] = n; g_nuniq[g] = 0; g_fallback[g] = 0; }
    if (n == 0) return;
    if (n > cap) { if (tid == 0) g_fallback[g] = 1; return; }
    // (1) stage the events, group equal hashes: slot per event, events per slot
    for (uint32_t i = tid; i < n; i += GRP_THREADS) {
        const uint4 *src = reinterpret_cast<const uint4 *>(part + e0 + i);
        const uint4 a = __ldcs(src), b = __ldcs(src + 1);  // read once
        const unsigned long long h = ((unsigned long long)a.y << 32) | a.x;
        S.rf[i] = ((uint64_t)a.w << 32) | a.z;
        S.p0[i] = ((uint64_t)b.y << 32) | b.x;
        S.p1[i] = ((uint64_t)b.w << 32) | b.z;
        uint32_t sl = (uint32_t)(h ^ (h >> 23)) & (GRP_SLOTS - 1);
        for (;;) {
            const unsigned long long prev = atomicCAS(&S.ht[sl], 0xFFFFFFFFFFFFFFFFull, h);
            if (prev == 0xFFFFFFFFFFFFFFFFull || prev == h) break;
            sl = (sl + 1) & (GRP_SLOTS - 1);
        }
        S.ev_slot[i] = (uint16_t)sl;
        atomicAdd(reinterpret_cast<unsigned int *>(S.scnt) + (sl >> 1), (sl & 1) ? 0x10000u : 1u);  // u16 counters, n <= 1024
    }
    __syncthreads();
    // exclusive scans over the slots: events per slot -> soff, occupied slots -> occ[]
    uint32_t nu;
    {
        constexpr int PER = GRP_SLOTS / GRP_THREADS;
        uint32_t loc[PER], tot = 0, oc = 0;
#pragma unroll
        for (int e = 0; e < PER; e++) { loc[e] = S.scnt[tid * PER + e]; tot += loc[e]; oc += loc[e] ? 1u : 0u; }
        uint32_t inc = tot, oinc = oc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, d), u = __shfl_up_sync(0xffffffffu, oinc, d);
            if (lane >= d) { inc += t; oinc += u; }
        }
        if (lane == 31) { S.wtot[wid] = inc; S.wocc[wid] = oinc; }
        __syncthreads();
        uint32_t base = inc - tot, obase = oinc - oc, ototal = 0;
#pragma unroll
        for (int w = 0; w < GRP_THREADS / 32; w++) { if (w < wid) { base += S.wtot[w]; obase += S.wocc[w]; } ototal += S.wocc[w]; }
        nu = ototal;
#pragma unroll
        for (int e = 0; e < PER; e++) {
            S.soff[tid * PER + e] = (uint16_t)base;
            base += loc[e];
            if (loc[e]) S.occ[obase++] = (uint16_t)(tid * PER + e);
        }
        if (tid == GRP_THREADS - 1) S.soff[GRP_SLOTS] = (uint16_t)base;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < PER; e++) S.scnt[tid * PER + e] = 0;  // reused as fill cursors
        __syncthreads();
    }
    for (uint32_t i = tid; i < n; i += GRP_THREADS) {
        const uint32_t sl = S.ev_slot[i];
        const unsigned int old = atomicAdd(reinterpret_cast<unsigned int *>(S.scnt) + (sl >> 1), (sl & 1) ? 0x10000u : 1u);
        const uint32_t k = (sl & 1) ? (old >> 16) : (old & 0xFFFFu);
        S.member[S.soff[sl] + k] = (uint16_t)i;
    }
    __syncthreads();
    // (2) replay dup_removal_lsh_full_exact (src/sketch.rs:690-731) per k-mer; the count lands in scnt.
    //     One thread per k-mer, registers only: a single pass over the k-mer's events keeps the four
    //     smallest (read index, event) keys; those four are replayed in order.  The dedup set at any
    //     point is simply "every pair key of the earlier paired events", so membership is a compare
    //     against the earlier events' keys and nothing has to be inserted.  Four counted events reach
    //     MAX_DEDUP_COUNT (src/constants.rs:14), after which every event counts.  Only a k-mer that
    //     has more than four events AND a duplicate among the first four needs more; it is queued
    //     for the warp-cooperative path below.
    //     The thread also counts its k-mer into its bucket (local index) for the output ordering.
    const uint32_t lb_n = S.lb_n;
    uint16_t *lb_cnt = S.ev_slot, *lb_off = S.ev_slot + GRP_LB;  // ev_slot is dead: 2 x GRP_LB u16
    if (lb_n <= (uint32_t)GRP_LB) {
        for (int i = tid; i < GRP_LB; i += GRP_THREADS) { lb_cnt[i] = 0; }
        __syncthreads();
    }
    uint32_t my_dups = 0;
    for (uint32_t u = tid; u < nu; u += GRP_THREADS) {
        const uint32_t sl = S.occ[u];
        const uint32_t a0 = S.soff[sl], len = S.soff[sl + 1] - a0;
        if (lb_n <= (uint32_t)GRP_LB) {
            const uint32_t lb = bucket_of(S.ht[sl], Mb, nbk) - bf;
            atomicAdd(reinterpret_cast<unsigned int *>(lb_cnt) + (lb >> 1), (lb & 1) ? 0x10000u : 1u);
        }
        uint32_t c = 0;
        if (no_dedup || len == 1) {
            c = len;  // a first occurrence is never a duplicate (c == 0)
        } else {
            constexpr uint64_t INF = 0xFFFFFFFFFFFFFFFFull;
            uint64_t k0 = INF, k1 = INF, k2 = INF, k3 = INF;  // (recflag << 10 | event) ascending
            for (uint32_t e = 0; e < len; e++) {
                const uint32_t m = S.member[a0 + e];
                uint64_t key = (S.rf[m] << 10) | m;
                if (key < k3) {
                    k3 = key;
                    if (k3 < k2) { const uint64_t t = k2; k2 = k3; k3 = t; }
                    if (k2 < k1) { const uint64_t t = k1; k1 = k2; k2 = t; }
                    if (k1 < k0) { const uint64_t t = k0; k0 = k1; k1 = t; }
                }
            }
            const uint64_t ks[4] = {k0, k1, k2, k3};
            uint64_t A[4], B[4];
            bool paired[4];
            uint32_t dups = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                paired[j] = false;
                A[j] = B[j] = 0;
                if (ks[j] == INF) continue;  // len < 4
                const uint32_t m = (uint32_t)ks[j] & 1023u;
                if ((ks[j] >> 10) & NO_PAIR) { c++; continue; }
                const uint64_t ka = S.p0[m], kb = S.p1[m];
                bool found = (kb == ka);
#pragma unroll
                for (int i = 0; i < j; i++)
                    found |= paired[i] && (ka == A[i] || ka == B[i] || kb == A[i] || kb == B[i]);
                paired[j] = true; A[j] = ka; B[j] = kb;
                // found on the SECOND key only because it equals the first does not make a duplicate
                // of an earlier read, but sketch.rs:704-722 still returns true when cur > 0
                if (found && c > 0) dups++; else c++;
            }
            if (len > 4u) {
                if (c >= 4u) c += len - 4u;
                else {  // a duplicate among the first four: replay the whole k-mer cooperatively
                    const uint32_t li = atomicAdd(&S.nlong, 1u);
                    S.longs[li] = (uint16_t)sl;
                    continue;
                }
            }
            my_dups += dups;
        }
        S.scnt[sl] = (uint16_t)c;  // c <= len <= GRP_CAP
    }
    __syncthreads();
    for (uint32_t li = wid; li < S.nlong; li += GRP_THREADS / 32) {
        const uint32_t sl = S.longs[li];
        const uint32_t a0 = S.soff[sl], len = S.soff[sl + 1] - a0;
        uint32_t c = 0, done = 0, nset = 0;
        uint64_t last_key = 0, dset = 0;  // lane q keeps dedup-set entry q in `dset` (32 entries, then the group falls back)
        uint32_t last_m = 0;
        bool first = true;
        while (done < len) {
            if (c >= 4u) { c += len - done; break; }  // MAX_DEDUP_COUNT
            if (done >= (uint32_t)GRP_SELECT_STEPS) { if (lane == 0) S.overflow = 1; break; }  // duplicate-heavy
            // next event after (last_key, last_m) in (read index, event index) order
            uint64_t best = 0xFFFFFFFFFFFFFFFFull;
            uint32_t bm = 0xFFFFFFFFu;
            for (uint32_t e = lane; e < len; e += 32) {
                const uint32_t m = S.member[a0 + e];
                const uint64_t key = S.rf[m];
                const bool after = first || key > last_key || (key == last_key && m > last_m);
                if (after && (key < best || (key == best && m < bm))) { best = key; bm = m; }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const uint64_t ok = __shfl_xor_sync(0xffffffffu, best, d);
                const uint32_t om = __shfl_xor_sync(0xffffffffu, bm, d);
                if (ok < best || (ok == best && om < bm)) { best = ok; bm = om; }
            }
            first = false; last_key = best; last_m = bm; done++;
            if (best & NO_PAIR) { c++; continue; }
            const uint64_t ka = S.p0[bm], kb = S.p1[bm];
            const bool in_a = __any_sync(0xffffffffu, lane < (int)nset && dset == ka);
            bool ret = in_a && c > 0;
            if (!in_a) { if (lane == (int)nset) dset = ka; nset++; }
            const bool in_b = __any_sync(0xffffffffu, lane < (int)nset && lane < 32 && dset == kb);
            if (in_b) ret = ret || c > 0;
            else { if (lane == (int)nset) dset = kb; nset++; }
            if (nset > 32u) { if (lane == 0) S.overflow = 1; break; }
            if (ret) { if (lane == 0) my_dups++; } else c++;
        }
        if (lane == 0) S.scnt[sl] = (uint16_t)c;
    }
    if (my_dups) atomicAdd(&S.dups, my_dups);
    __syncthreads();
    if (S.overflow) { if (tid == 0) g_fallback[g] = 1; return; }
    // (3) output order.  Bucket-rank path: lb_cnt holds the k-mers per bucket (counted in (2)).
    if (lb_n <= (uint32_t)GRP_LB) {
        uint16_t *list = S.member;  // event grouping is dead: k-mer slots ordered by bucket
        {   // exclusive scan of lb_cnt (GRP_LB = 2 per thread) -> lb_off, counters reset as cursors
            constexpr int PER = GRP_LB / GRP_THREADS;
            uint32_t loc[PER], tot = 0;
#pragma unroll
            for (int e = 0; e < PER; e++) { loc[e] = lb_cnt[tid * PER + e]; tot += loc[e]; }
            uint32_t inc = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
            if (lane == 31) S.wtot[wid] = inc;
            __syncthreads();
            uint32_t base = inc - tot;
#pragma unroll
            for (int w = 0; w < GRP_THREADS / 32; w++) if (w < wid) base += S.wtot[w];
#pragma unroll
            for (int e = 0; e < PER; e++) { lb_off[tid * PER + e] = (uint16_t)base; base += loc[e]; lb_cnt[tid * PER + e] = 0; }
            __syncthreads();
        }
        for (uint32_t u = tid; u < nu; u += GRP_THREADS) {
            const uint32_t sl = S.occ[u];
            const uint32_t lb = bucket_of(S.ht[sl], Mb, nbk) - bf;
            const unsigned int old = atomicAdd(reinterpret_cast<unsigned int *>(lb_cnt) + (lb >> 1), (lb & 1) ? 0x10000u : 1u);
            list[lb_off[lb] + ((lb & 1) ? (old >> 16) : (old & 0xFFFFu))] = (uint16_t)sl;
        }
        __syncthreads();
        for (uint32_t u = tid; u < nu; u += GRP_THREADS) {
            const uint32_t sl = S.occ[u];
            const unsigned long long h = S.ht[sl];
            const uint32_t lb = bucket_of(h, Mb, nbk) - bf;
            const uint32_t q0 = lb_off[lb], q1 = q0 + ((lb & 1) ? (reinterpret_cast<unsigned int *>(lb_cnt)[lb >> 1] >> 16)
                                                                : (reinterpret_cast<unsigned int *>(lb_cnt)[lb >> 1] & 0xFFFFu));
            uint32_t r = 0;
            for (uint32_t q = q0; q < q1; q++) r += (S.ht[list[q]] < h) ? 1u : 0u;
            st_hash[e0 + q0 + r] = h;
            st_cnt[e0 + q0 + r] = S.scnt[sl];
        }
        if (tid == 0) { g_nuniq[g] = nu; if (S.dups) atomicAdd(n_dup, (unsigned long long)S.dups); }
        return;
    }
    // Bitonic path: unique pairs into the (now dead) rf / p0 arrays, sorted by hash
    uint64_t *uh = S.rf;
    uint32_t *uc = reinterpret_cast<uint32_t *>(S.p0);
    uint32_t P = 32;
    while (P < nu) P <<= 1;
    for (uint32_t i = tid; i < P; i += GRP_THREADS) {
        if (i < nu) { const uint32_t sl = S.occ[i]; uh[i] = S.ht[sl]; uc[i] = S.scnt[sl]; }
        else { uh[i] = 0xFFFFFFFFFFFFFFFFull; uc[i] = 0; }
    }
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = tid; t < (P >> 1); t += GRP_THREADS) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const uint32_t l = i | j;
                const bool up = (i & k) == 0;
                const uint64_t ha = uh[i], hb = uh[l];
                if ((ha > hb) == up) {
                    uh[i] = hb; uh[l] = ha;
                    const uint32_t x = uc[i]; uc[i] = uc[l]; uc[l] = x;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = tid; i < nu; i += GRP_THREADS) { st_hash[e0 + i] = uh[i]; st_cnt[e0 + i] = uc[i]; }
    if (tid == 0) { g_nuniq[g] = nu; if (S.dups) atomicAdd(n_dup, (unsigned long long)S.dups); }
}

// staged pairs of group g live at [e0, e0 + nuniq); dense destination starts at uoff[g]
__global__ void k_compact_uniq(const uint64_t *__restrict__ st_hash, const uint32_t *__restrict__ st_cnt,
                               const uint32_t *__restrict__ g_e0, const uint32_t *__restrict__ g_nuniq,
                               const uint32_t *__restrict__ uoff, const uint8_t *__restrict__ g_fallback,
                               const uint32_t *__restrict__ g_src, const uint64_t *__restrict__ f_hash,
                               const uint32_t *__restrict__ f_cnt, uint64_t *__restrict__ out_hash,
                               uint32_t *__restrict__ out_cnt) {
    const uint32_t g = blockIdx.x, nu = g_nuniq[g], u0 = uoff[g];
    const bool fb = g_fallback[g] != 0;
    const uint64_t *sh = fb ? f_hash + g_src[g] : st_hash + g_e0[g];
    const uint32_t *sc = fb ? f_cnt + g_src[g] : st_cnt + g_e0[g];
    for (uint32_t i = threadIdx.x; i < nu; i += blockDim.x) {
        out_hash[u0 + i] = sh[i];
        out_cnt[u0 + i] = sc[i];
    }
}

// The generic path returns the fallback groups' unique pairs as ONE list sorted by hash.  Groups own
// disjoint, increasing hash ranges (bucket = mulhi(hash, Mb) is monotone), so group g's slice is
// [lower_bound(hash >= first hash of bucket bf), lower_bound(hash >= first hash of bucket be)).
__global__ void k_fallback_place(const uint8_t *__restrict__ g_fallback, const uint32_t *__restrict__ g_bf,
                                 const uint32_t *__restrict__ g_be, uint32_t ng, uint32_t nbk, uint64_t Mb,
                                 const uint64_t *__restrict__ f_hash, uint64_t fu, uint32_t *__restrict__ g_nuniq,
                                 uint32_t *__restrict__ g_src) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ng || !g_fallback[g]) return;
    auto first_hash_of = [&](uint32_t b) -> uint64_t {  // smallest h with mulhi(h, Mb) >= b
        const unsigned __int128 num = ((unsigned __int128)b << 64) + (Mb - 1);
        return (uint64_t)(num / Mb);
    };
    auto lb = [&](uint64_t v) -> uint64_t {
        uint64_t lo = 0, hi = fu;
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (f_hash[mid] < v) lo = mid + 1; else hi = mid; }
        return lo;
    };
    const uint32_t bf = g_bf[g], be = g_be[g];
    const uint64_t a = bf == 0 ? 0 : lb(first_hash_of(bf));
    const uint64_t b = be >= nbk ? fu : lb(first_hash_of(be));
    g_src[g] = (uint32_t)a;
    g_nuniq[g] = (uint32_t)(b - a);
}

// events of fallback groups -> compact SoA arrays for the generic path
__global__ void k_gather_fallback(const EventRec *__restrict__ part, const uint32_t *__restrict__ g_e0,
                                  const uint32_t *__restrict__ g_n, const uint8_t *__restrict__ g_fallback,
                                  const uint32_t *__restrict__ foff, uint64_t *__restrict__ hash,
                                  uint64_t *__restrict__ recflag, uint64_t *__restrict__ p0, uint64_t *__restrict__ p1) {
    const uint32_t g = blockIdx.x;
    if (!g_fallback[g]) return;
    const uint32_t e0 = g_e0[g], n = g_n[g], f0 = foff[g];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const EventRec r = part[e0 + i];
        hash[f0 + i] = r.hash; recflag[f0 + i] = r.recflag; p0[f0 + i] = r.p0; p1[f0 + i] = r.p1;
    }
}

__global__ void k_fallback_sizes(const uint32_t *__restrict__ g_n, const uint8_t *__restrict__ g_fallback, uint32_t ng,
                                 uint32_t *__restrict__ fsz) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < ng) fsz[g] = g_fallback[g] ? g_n[g] : 0u;
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
    DevBuf<EventRec> b_ev;  // event array (a scratch block of the ctx cache)
    EventRec *ev = nullptr;
    // Post-pass buckets: fixed before the first batch from the expected number of events, so that
    // k_seed can fill the bucket histogram while it flushes the events.
    uint64_t expect_bases = 0, expect_reads = 0;
    uint32_t nbk = 0;
    uint64_t Mb = 0;
    DevBuf<uint32_t> cnt;
    bool hist_valid = true;  // false after a capacity retry (partial counts of the failed attempt)

    int plan_buckets() {
        if (nbk) return SYL_OK;
        const uint64_t win = expect_bases > expect_reads * (uint64_t)(k - 1) ? expect_bases - expect_reads * (uint64_t)(k - 1) : 0;
        const uint64_t n_exp = win / c;
        nbk = 4096;
        while (nbk < n_exp / 32 && nbk < (1u << 22)) nbk <<= 1;
        const uint64_t thr = fmh_threshold(c);
        unsigned __int128 mb = ((unsigned __int128)nbk << 64) / ((unsigned __int128)thr + 1);
        Mb = mb > (unsigned __int128)UINT64_MAX ? UINT64_MAX : (uint64_t)mb;
        SYL_TRY(cnt.alloc(nbk, ctx->stream));
        SYL_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)nbk * 4, ctx->stream));
        return SYL_OK;
    }

    int reserve(uint64_t need) {
        if (need <= cap) return SYL_OK;
        const uint64_t ncap = std::max<uint64_t>(need, cap * 2);
        cudaStream_t st = ctx->stream;
        DevBuf<EventRec> ne;
        SYL_TRY(ne.alloc(ncap, st));
        if (n_events) SYL_CUDA(cudaMemcpyAsync(ne.p, ev, n_events * sizeof(EventRec), cudaMemcpyDeviceToDevice, st));
        b_ev.swap(ne);  // the old block goes back to the cache
        ev = b_ev.p;
        cap = ncap;
        return SYL_OK;
    }

    // one batch of reads, device resident; read indices continue from the previous batch.
    // k_seed appends the batch's events (hash, read, pair keys) straight to the event array.
    int add(const uint8_t *d_bases, uint64_t nb, const uint64_t *d_off, uint64_t off_bias, uint64_t nr) {
        cudaStream_t st = ctx->stream;
        if (nr == 0) return SYL_OK;
        uint64_t scap = nb / c + nb / (4 * c) + 65536;
        if (scap > nb) scap = nb + 16;
        uint64_t n = 0, npend = 0;
        DevBuf<uint32_t> pend;
        SYL_TRY(plan_buckets());
        for (;;) {
            if (scap >= 0xFFFFFFFFull) { set_error("more than 2^32-2 survivor events in one batch"); return SYL_ERR_ARG; }
            SYL_TRY(reserve(n_events + scap));
            SYL_TRY(pend.alloc(scap, st));
            int rc = seed_device_ex(ctx, d_bases, nb, d_off, off_bias, nr, k, c, sem, /*with_pos=*/0, ev + n_events, scap, &n,
                                    /*emit_events=*/1, n_reads, no_dedup, pend.p, &npend, cnt.p, Mb, nbk);
            if (rc == SYL_ERR_CAPACITY) { scap = n + 16; hist_valid = false; continue; }
            if (rc != SYL_OK) return rc;
            break;
        }
        if (npend) {
            const unsigned grid = (unsigned)std::min<uint64_t>(nblk(npend, EV_THREADS), 4096);
            k_events_fix<<<grid, EV_THREADS, 0, st>>>(ev + n_events, pend.p, reinterpret_cast<const unsigned long long *>(ctx->d_counters + 1),
                                                     d_bases, d_off, off_bias, n_reads);
            ctx->launches++;
            SYL_CUDA(cudaGetLastError());
        }
        n_events += n;
        n_reads += nr;
        n_bases += nb;
        return SYL_OK;
    }

    // Generic path: two stable LSD radix sorts (read index, then hash) + run-length encode +
    // k_dedup.  Handles any segment length in linear time; used for the groups the primary
    // path hands over (and for everything when SYL_SAMPLE_POSTPASS=sort).
    int dedup_sorted(const uint64_t *ev_hash, const uint64_t *ev_recflag, const uint64_t *ev_p0, const uint64_t *ev_p1,
                     uint64_t N, DevBuf<uint64_t> &uniq, DevBuf<uint32_t> &count, uint64_t *U_out, uint64_t *ndup_out) {
        cudaStream_t st = ctx->stream;
        *U_out = 0;
        *ndup_out = 0;
        if (N == 0) return SYL_OK;
        DevBuf<uint32_t> idx_a, idx_b;
        DevBuf<uint64_t> key_a, key_b;
        SYL_TRY(idx_a.alloc(N, st)); SYL_TRY(idx_b.alloc(N, st));
        SYL_TRY(key_a.alloc(N, st)); SYL_TRY(key_b.alloc(N, st));
        k_iota<<<nblk(N, 256), 256, 0, st>>>(idx_a.p, N);
        ctx->launches++;
        const int hash_bits = bits_for(fmh_threshold(c));
        DevBuf<uint8_t> tmp;
        size_t tmp_bytes = 0, t2 = 0;
        uint32_t *ord = idx_a.p;  // final event order
        uint64_t *hs = key_a.p;   // hashes in final order
        if (!no_dedup) {
            const int rec_bits = bits_for((n_reads << 1) | 1);
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ev_recflag, key_b.p, idx_a.p, idx_b.p, N, 0, rec_bits, st);
            cub::DeviceRadixSort::SortPairs(nullptr, t2, key_a.p, key_b.p, idx_b.p, idx_a.p, N, 0, hash_bits, st);
            tmp_bytes = std::max(tmp_bytes, t2);
            SYL_TRY(tmp.alloc(tmp_bytes, st));
            SYL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, ev_recflag, key_b.p, idx_a.p, idx_b.p, N, 0, rec_bits, st));
            k_gather<uint64_t><<<nblk(N, 256), 256, 0, st>>>(ev_hash, idx_b.p, key_a.p, N);
            SYL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, key_a.p, key_b.p, idx_b.p, idx_a.p, N, 0, hash_bits, st));
            ctx->launches += 3;
            hs = key_b.p;
            ord = idx_a.p;
        } else {
            cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, ev_hash, key_b.p, N, 0, hash_bits, st);
            SYL_TRY(tmp.alloc(tmp_bytes, st));
            SYL_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, ev_hash, key_b.p, N, 0, hash_bits, st));
            ctx->launches += 1;
            hs = key_b.p;
        }
        DevBuf<uint64_t> seg_off;
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
        ctx->launches += 1;
        SYL_TRY(count.alloc(std::max<uint64_t>(U, 1), st));
        if (no_dedup) {
            k_copy_len<<<nblk(U, 256), 256, 0, st>>>(seg_len.p, count.p, U);
            ctx->launches++;
        } else {
            tb = std::max(rle_bytes, scan_bytes);
            SYL_CUDA(cub::DeviceScan::ExclusiveSum(tmp2.p, tb, seg_len.p, seg_off.p, U, st));
            DevBuf<uint64_t> set;
            SYL_TRY(set.alloc(2 * N, st));
            unsigned long long *d_ndup = reinterpret_cast<unsigned long long *>(ctx->d_counters + 2);
            SYL_CUDA(cudaMemsetAsync(d_ndup, 0, 8, st));
            k_dedup<<<nblk(U, 128), 128, 0, st>>>(seg_off.p, seg_len.p, U, ord, ev_recflag, ev_p0, ev_p1, set.p, count.p, d_ndup);
            ctx->launches += 2;
            SYL_CUDA(cudaGetLastError());
            SYL_CUDA(cudaMemcpyAsync(ctx->h_counters + 2, d_ndup, 8, cudaMemcpyDeviceToHost, st));
            SYL_CUDA(cudaStreamSynchronize(st));
            *ndup_out = ctx->h_counters[2];
        }
        *U_out = U;
        return SYL_OK;
    }

    int finish(syl_sample **out) {
        cudaStream_t st = ctx->stream;
        syl_sample *s = new (std::nothrow) syl_sample();
        if (!s) return SYL_ERR_OOM;
        s->device = ctx->device;
        s->owner = ctx;
        s->stream = ctx->stream;
        s->k = k;
        s->c = c;
        s->mean_read_length = n_reads ? (double)n_bases / (double)n_reads : 0.;
        const uint64_t N = n_events;
        if (N == 0) { *out = s; return SYL_OK; }
        if (N >= 0xFFFFFFFFull) { delete s; set_error("more than 2^32-2 survivor events in one sample"); return SYL_ERR_ARG; }
        static const bool force_sort = []() { const char *e = getenv("SYL_SAMPLE_POSTPASS"); return e && std::string(e) == "sort"; }();
        static const uint32_t grp_cap = []() { const char *e = getenv("SYL_GROUP_CAP"); int v = e ? atoi(e) : GRP_CAP; return (uint32_t)std::min(std::max(v, 32), GRP_CAP); }();
        auto fail = [&](int rc) { syl_sample_free(s); return rc; };
        int rc;
        if (force_sort) {
            DevBuf<uint64_t> uq;
            DevBuf<uint32_t> ct;
            uint64_t U = 0, nd = 0;
            DevBuf<uint64_t> e_h, e_rf, e_p0, e_p1;
            if ((rc = e_h.alloc(N, st)) || (rc = e_rf.alloc(N, st)) || (rc = e_p0.alloc(N, st)) || (rc = e_p1.alloc(N, st))) return fail(rc);
            k_unpack_events<<<nblk(N, 256), 256, 0, st>>>(ev, N, e_h.p, e_rf.p, e_p0.p, e_p1.p);
            ctx->launches++;
            if ((rc = dedup_sorted(e_h.p, e_rf.p, e_p0.p, e_p1.p, N, uq, ct, &U, &nd)) != SYL_OK) return fail(rc);
            SYL_TRY(hblock_alloc(ctx, (void **)&s->hash, std::max<uint64_t>(U, 1) * 8));
            SYL_TRY(hblock_alloc(ctx, (void **)&s->count, std::max<uint64_t>(U, 1) * 4));
            SYL_CUDA(cudaMemcpyAsync(s->hash, uq.p, U * 8, cudaMemcpyDeviceToDevice, st));
            SYL_CUDA(cudaMemcpyAsync(s->count, ct.p, U * 4, cudaMemcpyDeviceToDevice, st));
            SYL_CUDA(cudaStreamSynchronize(st));
            s->n = U;
            s->num_dup_removed = nd;
            *out = s;
            return SYL_OK;
        }
        // ---- primary path: bucket partition + CTA-local sort/dedup (nbk, Mb, cnt: plan_buckets)
        const uint32_t ng = (uint32_t)(N / GRP_T + 1);
        DevBuf<uint32_t> boff, cursor, st_cnt, g_nuniq, g_e0, g_n, uoff, fsz, foff, g_bf, g_be, g_src;
        DevBuf<uint64_t> st_hash;
        DevBuf<uint8_t> g_fb;
        DevBuf<EventRec> part;
        if ((rc = boff.alloc((uint64_t)nbk + 1, st)) || (rc = cursor.alloc(nbk, st)) ||
            (rc = part.alloc(N, st)) || (rc = st_hash.alloc(N, st)) || (rc = st_cnt.alloc(N, st)) ||
            (rc = g_nuniq.alloc(ng, st)) || (rc = g_e0.alloc(ng, st)) || (rc = g_n.alloc(ng, st)) ||
            (rc = g_fb.alloc(ng, st)) || (rc = uoff.alloc((uint64_t)ng + 1, st)) || (rc = fsz.alloc(ng, st)) ||
            (rc = foff.alloc((uint64_t)ng + 1, st)) || (rc = g_bf.alloc(ng, st)) || (rc = g_be.alloc(ng, st)) ||
            (rc = g_src.alloc(ng, st)))
            return fail(rc);
        unsigned long long *d_ndup = reinterpret_cast<unsigned long long *>(ctx->d_counters + 2);
        SYL_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)nbk * 4, st));
        SYL_CUDA(cudaMemsetAsync(d_ndup, 0, 8, st));
        if (!hist_valid) {  // recount: a retried batch left partial counts behind
            SYL_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)nbk * 4, st));
            k_bucket_hist<<<nblk(N, 256), 256, 0, st>>>(ev, N, Mb, nbk, cnt.p);
            ctx->launches++;
        }
        {   // boff = exclusive scan of cnt (nbk >= 4096 entries): local scans, scan of block totals, add back
            const uint32_t nblk1 = nbk / 1024;
            DevBuf<uint32_t> btot, boff2;
            if ((rc = btot.alloc(nblk1, st)) || (rc = boff2.alloc((uint64_t)nblk1 + 1, st))) return fail(rc);
            k_scan_local<<<nblk1, 1024, 0, st>>>(cnt.p, nbk, boff.p, btot.p);
            k_scan_u32<<<1, 1024, 0, st>>>(btot.p, nblk1, boff2.p);
            k_scan_add<<<nblk1, 1024, 0, st>>>(boff.p, nbk, boff2.p);
            ctx->launches += 2;
        }
        k_scatter_events<<<nblk(N, 256), 256, 0, st>>>(ev, N, Mb, nbk, boff.p, cursor.p, part.p);
        SYL_CUDA(cudaFuncSetAttribute(k_group_dedup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GroupSmem)));
        k_group_ranges<<<nblk(ng, 256), 256, 0, st>>>(boff.p, nbk, ng, g_bf.p, g_be.p);
        {
            KernelTimer kt(ctx, SYL_KERNEL_GROUP_DEDUP);
            k_group_dedup<<<ng, GRP_THREADS, sizeof(GroupSmem), st>>>(part.p, boff.p, g_bf.p, g_be.p, grp_cap, Mb, nbk, no_dedup, st_hash.p, st_cnt.p,
                                                                       g_nuniq.p, g_e0.p, g_n.p, g_fb.p, d_ndup);
        }
        k_scan_u32<<<1, 1024, 0, st>>>(g_nuniq.p, ng, uoff.p);
        k_fallback_sizes<<<nblk(ng, 256), 256, 0, st>>>(g_n.p, g_fb.p, ng, fsz.p);
        k_scan_u32<<<1, 1024, 0, st>>>(fsz.p, ng, foff.p);
        ctx->launches += 6;
        SYL_CUDA(cudaGetLastError());
        SYL_CUDA(cudaMemcpyAsync(ctx->h_counters + 1, uoff.p + ng, 4, cudaMemcpyDeviceToHost, st));
        SYL_CUDA(cudaMemcpyAsync(ctx->h_counters + 3, foff.p + ng, 4, cudaMemcpyDeviceToHost, st));
        SYL_CUDA(cudaMemcpyAsync(ctx->h_counters + 2, d_ndup, 8, cudaMemcpyDeviceToHost, st));
        SYL_CUDA(cudaStreamSynchronize(st));
        const uint64_t U1 = (uint32_t)ctx->h_counters[1], NF = (uint32_t)ctx->h_counters[3];
        uint64_t ndup = ctx->h_counters[2];
        static const bool dbg = getenv("SYL_DEBUG_TIMING") != nullptr;
        if (dbg) fprintf(stderr, "[sample post-pass] events %llu buckets %u groups %u: in-kernel uniques %llu, events handed to the generic path %llu (%.1f %%)\n",
                         (unsigned long long)N, nbk, ng, (unsigned long long)U1, (unsigned long long)NF, 100.0 * NF / N);
        // fallback groups through the generic path
        DevBuf<uint64_t> fh, f_rf, f_p0, f_p1, f_uq;
        DevBuf<uint32_t> f_ct;
        uint64_t U2 = 0, nd2 = 0;
        if (NF) {
            if ((rc = fh.alloc(NF, st)) || (rc = f_rf.alloc(NF, st)) || (rc = f_p0.alloc(NF, st)) || (rc = f_p1.alloc(NF, st)))
                return fail(rc);
            k_gather_fallback<<<ng, 256, 0, st>>>(part.p, g_e0.p, g_n.p, g_fb.p, foff.p, fh.p, f_rf.p, f_p0.p, f_p1.p);
            ctx->launches++;
            if ((rc = dedup_sorted(fh.p, f_rf.p, f_p0.p, f_p1.p, NF, f_uq, f_ct, &U2, &nd2)) != SYL_OK) return fail(rc);
            ndup += nd2;
        }
        const uint64_t U = U1 + U2;
        SYL_TRY(hblock_alloc(ctx, (void **)&s->hash, std::max<uint64_t>(U, 1) * 8));
        SYL_TRY(hblock_alloc(ctx, (void **)&s->count, std::max<uint64_t>(U, 1) * 4));
        if (U2) {  // slot the generic path's pairs into their groups' positions and redo the output offsets
            k_fallback_place<<<nblk(ng, 128), 128, 0, st>>>(g_fb.p, g_bf.p, g_be.p, ng, nbk, Mb, f_uq.p, U2, g_nuniq.p, g_src.p);
            k_scan_u32<<<1, 1024, 0, st>>>(g_nuniq.p, ng, uoff.p);
            ctx->launches += 2;
        }
        k_compact_uniq<<<ng, 128, 0, st>>>(st_hash.p, st_cnt.p, g_e0.p, g_nuniq.p, uoff.p, g_fb.p, g_src.p, f_uq.p, f_ct.p,
                                            s->hash, s->count);
        ctx->launches++;
        SYL_CUDA(cudaGetLastError());
        SYL_CUDA(cudaStreamSynchronize(st));
        s->n = U;
        s->num_dup_removed = ndup;
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
    syl::tl_ctx = ctx;
    SampleBuilder b{ctx, k, c, no_dedup, sem};
    b.expect_bases = n_bases;
    b.expect_reads = n_reads;
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
    syl::tl_ctx = ctx;
    cudaStream_t st = ctx->stream;
    syl_sample *s = new (std::nothrow) syl_sample();
    if (!s) return SYL_ERR_OOM;
    s->device = ctx->device; s->owner = ctx; s->stream = st; s->k = k; s->c = c; s->n = n;
    SYL_TRY(hblock_alloc(ctx, (void **)&s->hash, std::max<uint64_t>(n, 1) * 8));
    SYL_TRY(hblock_alloc(ctx, (void **)&s->count, std::max<uint64_t>(n, 1) * 4));
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
    syl::tl_ctx = ctx;
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
    hblock_free(s->owner, s->hash);   // back into the owning ctx's block cache
    hblock_free(s->owner, s->count);
    delete s;
}

}  // extern "C"
