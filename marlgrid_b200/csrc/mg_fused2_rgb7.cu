// mg_fused2_rgb7.cu -- instantiations of the specialised fused kernel: RGB observations, view size 7.
#include "mg_fused2.cuh"

namespace mg {
template int launch_fused2_ov<2, 7>(const KP&, cudaStream_t);
}
