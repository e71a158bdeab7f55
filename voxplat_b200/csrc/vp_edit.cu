// vp_edit.cu -- brush edits applied to the device-resident world.
//
// Replaces chunkset_edit_sphere (chunkset/edit.c:179-244) together with chunk_ws_write (:143-176) and
// shadow_place_update (shadow.h:77-89) for the device copy, so that an edit burst (BASELINE config C5) needs no
// host -> device voxel traffic at all:
//   - every cell u of the box [c-r-1, c+r+1) with |u - c| < r is written (cells with y < 2 are protected, edit.c:151;
//     the float distance test is exact in integers for any realistic radius);
//   - when a solid voxel is placed, the height map takes shadow_place_update for every such cell IN THE REFERENCE'S
//     ORDER (chunk list x,y,z outer; inside a chunk x outer, y, z inner): the rule is order dependent, but rows of
//     different z never interact, so one lane per z row replays its row sequentially;
//   - all chunks touched by the box become dirty (edit.c:241-242) -- the list is returned to the caller.
#include "vp_device.cuh"
using namespace vp;

namespace {

__global__ void __launch_bounds__(256)
k_edit_sphere(VpWorldDev w, uint8_t *__restrict__ vox_pool, uint16_t *__restrict__ shadow, int cx, int cy, int cz, int radius, int voxel,
              int own_z0, int own_z1)
{
	const int rb = w.rb, R = 1 << rb;
	const int X = 1 << (w.bits[0] + rb), Y = 1 << (w.bits[1] + rb), Z = 1 << (w.bits[2] + rb);
	const int lo[3] = { cx - radius - 1, cy - radius - 1, cz - radius - 1 };
	const int e = 2 * radius + 2;                                  // cells per axis: [c-r-1, c+r+1)
	const int r2 = radius * radius;
	// ---- voxel writes, one thread per cell ----
	for (int i = threadIdx.x; i < e * e * e; i += blockDim.x) {
		const int x = lo[0] + i / (e * e), y = lo[1] + (i / e) % e, z = lo[2] + i % e;
		if (x < 0 || y < 0 || z < 0 || x >= X || y >= Y || z >= Z) continue;          // chunk_ws_inside fails for every listed chunk
		const int dx = x - cx, dy = y - cy, dz = z - cz;
		if (dx * dx + dy * dy + dz * dz >= r2) continue;                                // glm_vec_distance(u, c) < radius (edit.c:221)
		if (y < 2) continue;                                                            // edit.c:151
		if ((z >> rb) < own_z0 || (z >> rb) >= own_z1) continue;                        // another GPU's slab
		const int slot = chunk_slot(w, x >> rb, y >> rb, z >> rb);
		if (slot < 0) continue;                                                         // cannot happen: the host gave every box chunk a slot
		vox_pool[((size_t)slot << (3 * rb)) + ((((size_t)(z & (R - 1)) << rb) | (size_t)(y & (R - 1))) << rb | (size_t)(x & (R - 1)))] = (uint8_t)voxel;
	}
	// ---- height map: one lane per z row, the reference's visiting order inside the row ----
	if (!voxel || threadIdx.x >= e) return;
	const int z = lo[2] + threadIdx.x;
	if (z < 0 || z >= Z) return;
	if ((unsigned)z < w.sh_z0 || (unsigned)z >= w.sh_z1) return;               // rows this device does not hold (a slab keeps [sh_z0, sh_z1))
	uint16_t *row = shadow + (size_t)(z - (int)w.sh_z0) * w.sh_w;
	const int hx = cx + radius + 1, hy = cy + radius + 1;           // exclusive upper corner of the box
	const int dz = z - cz;
	for (int gx = max(lo[0], 0) >> rb; gx <= min(hx, X - 1) >> rb && gx < (1 << w.bits[0]); gx++)
	for (int gy = max(lo[1], 0) >> rb; gy <= min(hy, Y - 1) >> rb && gy < (1 << w.bits[1]); gy++) {
		const int x0 = max(lo[0], gx << rb), x1 = min(hx, (gx + 1) << rb), y0 = max(lo[1], gy << rb), y1 = min(hy, (gy + 1) << rb);
		for (int x = x0; x < x1; x++) for (int y = y0; y < y1; y++) {
			const int dx = x - cx, dy = y - cy;
			if (dx * dx + dy * dy + dz * dz >= r2) continue;
			const uint32_t idx = (uint32_t)(x + y), lim = (uint32_t)y + 1u;           // shadow.h:45-63,77-89
			if (row[idx] >= lim || row[idx + 1] >= lim) continue;
			row[idx] = (uint16_t)y;
		}
	}
}

// chunkset_edit_raycast_until_solid (chunkset/edit.c:248-314): a DDA walk from `origin` along `vector` until a solid voxel,
// one thread per ray.  The float arithmetic restates the reference's expressions operation by operation -- IEEE divide,
// the sum of squares in the association the reference build uses (one multiply, two fused multiply-adds), float square
// root, unsigned <-> float conversions -- so that the walk takes the same side at every step, zero components of the
// vector included (their NaN / infinity distances compare the same way).  Coordinates that leave the world through a
// 0-face stick at 0xFFFFFFFF like the reference's unsigned conversion: such a ray never hits.  At most 4095 steps.
__device__ __forceinline__ uint32_t f2u_sat(float f)
{
	return (!(f > -1.0f) || f >= 4294967296.0f) ? 0xFFFFFFFFu : (uint32_t)f;
}

__global__ void __launch_bounds__(128)
k_raycast(VpWorldDev w, const uint8_t *__restrict__ vox_pool, uint32_t n, const float *__restrict__ origins, const float *__restrict__ vectors,
          uint32_t *__restrict__ coords, int8_t *__restrict__ normals, uint8_t *__restrict__ voxels)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int rb = w.rb, R = 1 << rb;
	const uint32_t X = 1u << (w.bits[0] + rb), Y = 1u << (w.bits[1] + rb), Z = 1u << (w.bits[2] + rb);
	const float o0 = origins[3 * i], o1 = origins[3 * i + 1], o2 = origins[3 * i + 2];
	const float v0 = vectors[3 * i], v1 = vectors[3 * i + 1], v2 = vectors[3 * i + 2];
	uint32_t c0 = (uint32_t)(int)o0, c1 = (uint32_t)(int)o1, c2 = (uint32_t)(int)o2;
	auto delta = [](float a, float b, float c) {      // t[i][0..2] = (a, b, c): fma(c, c, fma(a, a, b * b)), then sqrt
		return __fsqrt_rn(__fmaf_rn(c, c, __fmaf_rn(a, a, __fmul_rn(b, b))));
	};
	const float d0 = delta(__fdiv_rn(v0, v0), __fdiv_rn(v1, v0), __fdiv_rn(v2, v0));
	const float d1 = delta(__fdiv_rn(v0, v1), __fdiv_rn(v1, v1), __fdiv_rn(v2, v1));
	const float d2 = delta(__fdiv_rn(v0, v2), __fdiv_rn(v1, v2), __fdiv_rn(v2, v2));
	auto first = [](float o, uint32_t c, float v, float d, float &step) {
		if (0.0f > v) { step = -1.0f; return __fmul_rn(__fsub_rn(o, __uint2float_rn(c)), d); }
		step = 1.0f; return __fmul_rn(__fsub_rn(__fadd_rn(__uint2float_rn(c), 1.0f), o), d);
	};
	float s0, s1, s2;
	float n0 = first(o0, c0, v0, d0, s0), n1 = first(o1, c1, v1, d1, s1), n2 = first(o2, c2, v2, d2, s2);
	uint32_t hit = 0;
	int side = 0;
	for (int loops = 4095; loops > 0; loops--) {
		side = 0;
		if (n0 > n1) side = 1;
		if ((side ? n1 : n0) > n2) side = 2;
		if (side == 0) { n0 = __fadd_rn(n0, d0); c0 = f2u_sat(__fadd_rn(__uint2float_rn(c0), s0)); }
		else if (side == 1) { n1 = __fadd_rn(n1, d1); c1 = f2u_sat(__fadd_rn(__uint2float_rn(c1), s1)); }
		else { n2 = __fadd_rn(n2, d2); c2 = f2u_sat(__fadd_rn(__uint2float_rn(c2), s2)); }
		if (c0 >= X || c1 >= Y || c2 >= Z) continue;                          // chunkset_edit_read: 0 outside the world (edit.c:22-25)
		const int slot = chunk_slot(w, (int)(c0 >> rb), (int)(c1 >> rb), (int)(c2 >> rb));
		if (slot < 0) continue;                                               // null chunk (or a row another device holds)
		hit = vox_pool[((size_t)slot << (3 * rb)) + ((((size_t)(c2 & (R - 1)) << rb) | (size_t)(c1 & (R - 1))) << rb | (size_t)(c0 & (R - 1)))];
		if (hit) break;
	}
	coords[3 * i] = c0; coords[3 * i + 1] = c1; coords[3 * i + 2] = c2;
	voxels[i] = (uint8_t)hit;
	if (hit) {
		const float vs = side == 0 ? v0 : side == 1 ? v1 : v2;
		normals[3 * i + side] = (-vs > 0.0f) ? 1 : -1;                         // edit.c:304; the other two entries stay the caller's
	}
}

} // namespace

cudaError_t vp_launch_raycast(const VpWorldDev &w, const uint8_t *vox_pool, uint32_t n, const float *origins, const float *vectors,
                              uint32_t *coords, int8_t *normals, uint8_t *voxels, cudaStream_t s)
{
	if (!n) return cudaSuccess;
	k_raycast<<<(n + 127) / 128, 128, 0, s>>>(w, vox_pool, n, origins, vectors, coords, normals, voxels);
	return cudaGetLastError();
}

cudaError_t vp_launch_edit_sphere(const VpWorldDev &w, uint8_t *vox_pool, uint16_t *shadow, int cx, int cy, int cz, int radius, int voxel,
                                  int own_z0, int own_z1, cudaStream_t s)
{
	k_edit_sphere<<<1, 256, 0, s>>>(w, vox_pool, shadow, cx, cy, cz, radius, voxel, own_z0, own_z1);
	return cudaGetLastError();
}
