// Check-node kernels (ANMSA and AOMSA) for double messages, 2 frame(s) per lane.
#define QK_DEFINE_CN_LAUNCH
#include "run_batch.cuh"
namespace qkhost {
template int launch_cn_alg<double, 2, 4>(const qkdldpc_code *, bool, int, cudaStream_t, const qk::StepArgs<double> &);
template int launch_cn_alg<double, 2, 5>(const qkdldpc_code *, bool, int, cudaStream_t, const qk::StepArgs<double> &);
}
