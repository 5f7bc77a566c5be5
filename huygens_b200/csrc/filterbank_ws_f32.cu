// filterbank_ws_f32.cu — float instantiations of the warp-specialised exact kernel (filterbank_ws.inc).
#include "filterbank_ws.inc"
namespace hg {
template hg_status fb_launch_exact_ws<float>(bool, bool, int, int, const FbK1Params<float>&, int, cudaStream_t);
template bool fb_ws_describe<float>(int, int, int*, int*, int*, long long*);
}
