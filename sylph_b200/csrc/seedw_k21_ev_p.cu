// k_seed_w<K=21, read-sketch events, 2-bit packed input>: run lengths 24 / 30 / 32 (one translation unit per variant: parallel compile)
#include "seed_warp.cuh"
namespace syl {
SEEDW_DEFINE_KERNELS(seedw_kernels_k21_ev_p, 21, 1, true)
}
