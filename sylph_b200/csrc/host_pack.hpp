// host_pack.hpp — host side of the 2-bit ingest path: exact BYTE_TO_SEQ packing of ASCII
// bases into the forward-stream words the seeding kernel consumes, spread over a persistent pool of
// worker threads that fills a ring of pinned staging buffers while earlier chunks cross PCIe.
#pragma once
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace syl {

// words[w] = sum_j BYTE_TO_SEQ[bases[16w + j]] << (30 - 2j)  (src/types.rs:50-59; bases past n read as 0)
// n_bases need not be a multiple of 16; ceil(n / 16) words are written.  Single thread.
void pack2_range(const uint8_t *bases, uint64_t n_bases, uint32_t *words);

// A unit of packing work: one slice of one chunk.
struct PackItem {
    const uint8_t *src;   // ASCII bases of the slice (slice starts at a multiple of 16 bases inside its chunk)
    uint64_t n;           // bases in the slice
    uint32_t *dst;        // destination words
    const uint64_t *off_src;  // or: a slice of record offsets to rebase into u32 (src == nullptr)
    uint64_t off_base, off_n;
    uint32_t *off_dst;
    uint32_t chunk;       // chunk the slice belongs to
};

// Persistent worker pool.  run() hands out the items of `items` (ordered by chunk) to the workers;
// item i of chunk c may only start once gate(c) allows it (its staging buffer is free again), and
// chunk_done[c] counts the finished items of chunk c so that the caller can ship complete chunks
// while later ones are still being packed.
class PackPool {
public:
    explicit PackPool(int n_threads);
    ~PackPool();
    int threads() const { return (int)workers_.size(); }
    // start a job; returns immediately.  items / remaining stay owned by the caller and must outlive finish().
    void start(const std::vector<PackItem> *items, std::vector<std::atomic<uint32_t>> *remaining, std::atomic<int64_t> *gate_chunk);
    // block until chunk c is completely packed
    void wait_chunk(uint32_t c);
    // block until every item has been processed (must be called before the next start)
    void finish();

private:
    void worker();
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
    uint64_t generation_ = 0;
    const std::vector<PackItem> *items_ = nullptr;
    std::vector<std::atomic<uint32_t>> *remaining_ = nullptr;
    std::atomic<int64_t> *gate_ = nullptr;  // items of chunks <= *gate_ may run
    std::atomic<uint64_t> next_{0};
    std::atomic<int> active_{0};
};

int default_pack_threads();

}  // namespace syl
