#include "mg_common.cuh"
namespace mg {
int launch_fused2(const KP&, int, cudaStream_t) { return MG_E_UNSUPPORTED; }
}
