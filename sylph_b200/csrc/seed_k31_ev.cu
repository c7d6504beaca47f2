// k_seed<K=31, EMIT=1> for run lengths 24 / 30 / 32 (EMIT 0: 16-byte survivors, 1: 32-byte read-sketch events)
#include "seed_kernel.cuh"

namespace syl {
SEED_DEFINE_KERNELS(seed_kernels_k31_ev, 31, 1)
}  // namespace syl
