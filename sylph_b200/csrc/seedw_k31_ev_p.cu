// k_seed_w<K=31, read-sketch events, 2-bit packed input>: run lengths 24 / 30 / 32 (one translation unit per variant: parallel compile)
#include "seed_warp.cuh"
namespace syl {
SEEDW_DEFINE_KERNELS(seedw_kernels_k31_ev_p, 31, 1, true)
}
