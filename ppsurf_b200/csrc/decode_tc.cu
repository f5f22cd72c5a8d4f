// Tensor-core path of the decoder's global branch (path 1).  Placeholder until the tcgen05 kernel lands: it reports an
// error instead of silently falling back.
#include "common.cuh"

namespace pps {
size_t projection_tc_workspace(const pps_decoder_weights*, int64_t) { return 256; }
int projection_tc_impl(const pps_decoder_weights*, const float*, const float*, const int32_t*, int, int64_t, void*, size_t,
                       float*, cudaStream_t) {
    set_error("decoder path 1 (tcgen05) is not built into this library yet");
    return PPS_ERR_INVALID;
}
}  // namespace pps
