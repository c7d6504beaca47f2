// k_seed_w<K=21, read-sketch events, ASCII input>: run lengths 24 / 30 / 32 (one translation unit per variant: parallel compile)
#include "seed_warp.cuh"
namespace syl {
SEEDW_DEFINE_KERNELS(seedw_kernels_k21_ev, 21, 1, false)
}
