/* Shim for <cglm/cglm.h> (absent in this image).  The reference's chunk path uses exactly one
 * cglm symbol: glm_vec_distance (chunkset/edit.c:221).  TEST INFRASTRUCTURE ONLY. */
#ifndef VOXREF_SHIM_CGLM_H
#define VOXREF_SHIM_CGLM_H
#include <math.h>
static inline float glm_vec_distance(float *a, float *b)
{
	float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
	return sqrtf(dx * dx + dy * dy + dz * dz);
}
#endif
