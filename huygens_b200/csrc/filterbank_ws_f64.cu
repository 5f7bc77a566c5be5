// filterbank_ws_f64.cu — double instantiations of the warp-specialised exact kernel (filterbank_ws.inc).
#include "filterbank_ws.inc"
namespace hg {
template hg_status fb_launch_exact_ws<double>(bool, bool, int, int, const FbK1Params<double>&, int, cudaStream_t);
template bool fb_ws_describe<double>(int, int, int*, int*, int*, long long*);
}
