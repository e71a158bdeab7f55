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

} // namespace

cudaError_t vp_launch_edit_sphere(const VpWorldDev &w, uint8_t *vox_pool, uint16_t *shadow, int cx, int cy, int cz, int radius, int voxel,
                                  int own_z0, int own_z1, cudaStream_t s)
{
	k_edit_sphere<<<1, 256, 0, s>>>(w, vox_pool, shadow, cx, cy, cz, radius, voxel, own_z0, own_z1);
	return cudaGetLastError();
}
