// mg_fused2_enc7.cu -- instantiations of the specialised fused kernel: encoded observations, view size 7.
#include "mg_fused2.cuh"

namespace mg {
template int launch_fused2_ov<1, 7>(const KP&, cudaStream_t);
template int launch_fused2_ks<7>(const KP&, int, cudaStream_t);
}
