// mg_fused2_rgb5.cu -- instantiations of the specialised fused kernel: RGB observations, view size 5.
#include "mg_fused2.cuh"

namespace mg {
template int launch_fused2_ov<2, 5>(const KP&, cudaStream_t);
}
