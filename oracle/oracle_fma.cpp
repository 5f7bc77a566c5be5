// fma arithmetic mode: explicit std::fma in the fixed pattern of huygens_oracle.hpp; built with -mfma.
// TEST INFRASTRUCTURE ONLY — see huygens_oracle.hpp.
#define HGO_FMA 1
#define HGO_SUFFIX _fma
#include "oracle_capi.inc"
