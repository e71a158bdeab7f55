/* oracle/gfx_shim/cglm/cglm.h -- TEST INFRASTRUCTURE ONLY.  Just enough of cglm's types for the reference's gfx/vsplat.c
 * to compile headless (its draw code is never called by the harness; only gfx_update_svl runs). */
#pragma once
#include <math.h>
typedef float vec3[3];
typedef float vec4[4];
typedef vec4 mat4[4];
static inline float glm_vec_distance(float *a, float *b)
{
	float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
	return sqrtf(dx * dx + dy * dy + dz * dz);
}
