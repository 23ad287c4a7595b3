// mg_fused2_enc7h.cu -- instantiations of the specialised fused kernel: encoded observations with hide_item_types, view size 7.
#include "mg_fused2.cuh"

namespace mg {
template int launch_fused2_hide<7>(const KP&, cudaStream_t);
}
