// tcgen05/TMEM implicit-GEMM convolution - placeholder until the kernel lands.
#include "kernels.h"
namespace ledb {
bool conv_tc_eligible(const ConvArgs&) { return false; }
int launch_conv_tc(const ConvArgs&, cudaStream_t) { return fail(LEDB200_EINVAL, "tcgen05 conv not built"); }
}  // namespace ledb
