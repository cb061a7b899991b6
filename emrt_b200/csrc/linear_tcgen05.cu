// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
namespace emrt {
int linear_tcgen05(const emrt_linear_args*, cudaStream_t) {
  return set_error(EMRT_ERR_UNSUPPORTED, "tcgen05 linear not built yet");
}
}
