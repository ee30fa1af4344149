// epic_tc.cu — tcgen05 path (placeholder until the kernels land in this round)
#include "mmb_internal.h"
namespace mmb {
bool tc_supported(const MmbEpicDims*, int) { return false; }
int tc_build_image(EpicModel*, const float*) { return MMB_OK; }
int launch_epic_forward_tc(const EpicModel*, const float*, const uint8_t*, const uint8_t*, const float*, int, int, int,
                           float*, float*, float*, cudaStream_t) { return fail(MMB_EUNSUPPORTED, "tcgen05 path not built"); }
int launch_generate_tc(const EpicModel*, float*, uint8_t*, const uint8_t*, const float*, int, float, const float*,
                       uint64_t, uint64_t, int, int, cudaStream_t) { return fail(MMB_EUNSUPPORTED, "tcgen05 path not built"); }
}
