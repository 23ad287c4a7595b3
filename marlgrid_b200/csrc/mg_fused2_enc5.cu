// mg_fused2_enc5.cu -- instantiations of the specialised fused kernel: encoded observations, view size 5.
#include "mg_fused2.cuh"

namespace mg {
template int launch_fused2_ov<1, 5>(const KP&, cudaStream_t);
template int launch_fused2_ks<5>(const KP&, int, cudaStream_t);
}
