// Step loop, scheduler, prep and variable-node kernels for double messages, 2 frame(s) per lane (tile = 64 frames).
#define QK_DEFINE_RUN_BATCH
#include "run_batch.cuh"
namespace qkhost {
template int run_batch<double, 2>(qkdldpc_code *, const qkdldpc_params *, int64_t, const uint32_t *, const uint32_t *, const double *, int, const int32_t *, int, const int32_t *, int, uint32_t *, int32_t *, uint8_t *, unsigned long long *);
}
