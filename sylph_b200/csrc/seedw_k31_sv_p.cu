// k_seed_w<K=31, survivors, 2-bit packed input>: run lengths 24 / 30 / 32 (one translation unit per variant: parallel compile)
#include "seed_warp.cuh"
namespace syl {
SEEDW_DEFINE_KERNELS(seedw_kernels_k31_sv_p, 31, 0, true)
}
