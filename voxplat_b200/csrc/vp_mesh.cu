// placeholder until the mesh kernel lands (next commit)
#include "vp_internal.h"
cudaError_t vp_launch_mesh(const VpWorldDev &, const uint32_t *, uint32_t n, VpResultDev *, const uint32_t *, uint8_t *, VpArenaDev *, cudaStream_t)
{ return n ? cudaErrorNotSupported : cudaSuccess; }
