// k_seed_w<K=31, read-sketch events, ASCII input>: run lengths 24 / 30 / 32 (one translation unit per variant: parallel compile)
#include "seed_warp.cuh"
namespace syl {
SEEDW_DEFINE_KERNELS(seedw_kernels_k31_ev, 31, 1, false)
}
