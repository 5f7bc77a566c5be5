// strict arithmetic mode: no contraction, no -march (mirrors /root/reference/tests/build.sh:8 on x86-64).
// TEST INFRASTRUCTURE ONLY — see huygens_oracle.hpp.
#define HGO_FMA 0
#define HGO_SUFFIX _strict
#include "oracle_capi.inc"
