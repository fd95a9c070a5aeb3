// Check-node kernels (ANMSA and AOMSA) for float messages, 4 frame(s) per lane.
#define QK_DEFINE_CN_LAUNCH
#include "run_batch.cuh"
namespace qkhost {
template int launch_cn_alg<float, 4, 4>(const qkdldpc_code *, bool, int, cudaStream_t, const qk::StepArgs<float> &);
template int launch_cn_alg<float, 4, 5>(const qkdldpc_code *, bool, int, cudaStream_t, const qk::StepArgs<float> &);
}
