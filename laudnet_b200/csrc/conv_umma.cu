// Placeholder until the tcgen05 kernel lands: routes to the HMMA kernel.
#include "laud_common.cuh"
namespace laud {
int conv_forward_umma(const ConvArgs& a, cudaStream_t s) { return conv_forward_hmma(a, s); }
}
