// placeholder until the tcgen05 GEMM lands: reports "not handled" so callers fall back to SIMT.
#include "common.cuh"
namespace nt {
int gemm_tc(const GemmDesc&, int, bool* handled) { *handled = false; return 0; }
}
