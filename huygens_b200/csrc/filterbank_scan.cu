// filterbank_scan.cu — K2: chunked time-parallel scan of the biquad state recurrence (placeholder
// until the kernel lands; fails loudly rather than falling back).
#include "hg_common.cuh"
#include "filterbank_handle.cuh"

namespace hg {
template <typename T>
hg_status fb_process_scan(hg_fb*, const T*, T*, size_t, cudaStream_t) {
    return fail(HG_ERR_UNSUPPORTED, "scan mode not built yet");
}
template hg_status fb_process_scan<float>(hg_fb*, const float*, float*, size_t, cudaStream_t);
template hg_status fb_process_scan<double>(hg_fb*, const double*, double*, size_t, cudaStream_t);
}  // namespace hg
