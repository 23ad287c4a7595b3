// mg_fused2_enc5h.cu -- instantiations of the specialised fused kernel: encoded observations with hide_item_types, view size 5.
#include "mg_fused2.cuh"

namespace mg {
template int launch_fused2_hide<5>(const KP&, cudaStream_t);
}
