// k_seed_w<K=21, survivors, 2-bit packed input>: run lengths 24 / 30 / 32 (one translation unit per variant: parallel compile)
#include "seed_warp.cuh"
namespace syl {
SEEDW_DEFINE_KERNELS(seedw_kernels_k21_sv_p, 21, 0, true)
}
