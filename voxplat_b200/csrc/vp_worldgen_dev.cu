// vp_worldgen_dev.cu -- world generation on the device: the deterministic integer generator of
// vp_worldgen.c (same source: vp_worldgen_core.h) as CUDA kernels, so a world never crosses PCIe, plus the shadow-map
// build from resident voxels.
//
// The reference's generator (chunkset/gen.c:187-341) takes its arithmetic from FastNoise (un-vendored, un-pinned) and
// is non-deterministic under OpenMP, so it cannot be an oracle; the parity target of these kernels is the host generator
// vp_worldgen.c, byte for byte (tests/test_gpu_worldgen.py), which keeps the structure of gen.c (height field, colour
// bands, trees on a 10-voxel lattice, shadow_place_update in a fixed order).
//
//   k_gen_chunks   one CTA per chunk: ground columns (lx fastest across lanes: 32-byte stores), then the trees that can
//                  reach the chunk, applied in the host's (gz,gx) order by one warp because later writes win.
//   k_shadow_rows  one CTA per z row, one thread per y: shadow_place_update (shadow.h:77-89) is order dependent along x
//                  but, inside one column, every y reads only OLD entries (idx = x+y and x+y+1, written by y and y+1 of
//                  the same column) -- so a column is one parallel step between two barriers and the row lives in
//                  shared memory.  This is the "parallel-safe shadow-map build" of the survey; same result as the
//                  sequential rule in the host's fixed order (x ascending, then y ascending).
#include "vp_internal.h"
#include "vp_worldgen_core.h"

namespace {

constexpr int kGenThreads = 256;
constexpr int kMaxCells = 448;          // tree lattice cells that can reach a chunk: ((R + 32) / 10 + 3)^2 <= 19^2 for R = 128

struct TreeRec { int32_t x, y, z, ok; };

__global__ void __launch_bounds__(kGenThreads)
k_gen_chunks(const vpw_params P, const uint32_t *__restrict__ ids, uint8_t *__restrict__ out, uint32_t *__restrict__ solid)
{
	__shared__ TreeRec trees[kMaxCells];
	__shared__ int s_written;
	const int rb = P.root_bitw, R = 1 << rb;
	const uint32_t id = ids[blockIdx.x];
	const int32_t cx = (int32_t)(id & ((1u << P.bits[0]) - 1)), cy = (int32_t)((id >> P.bits[0]) & ((1u << P.bits[1]) - 1));
	const int32_t cz = (int32_t)(id >> (P.bits[0] + P.bits[1]));
	const int32_t ox = cx << rb, oy = cy << rb, oz = cz << rb;
	uint8_t *chunk = out + ((size_t)blockIdx.x << (3 * rb));
	const int tid = threadIdx.x;
	if (tid == 0) s_written = 0;

	// ground: one column per thread and step, x fastest across the lanes
	int any = 0;
	for (int col = tid; col < R * R; col += kGenThreads) {
		const int lz = col >> rb, lx = col & (R - 1);
		const int32_t h = column_height(&P, (uint32_t)(ox + lx), (uint32_t)(oz + lz));
		int32_t top = h - oy; if (top > R) top = R;
		const int32_t ls = h - 1 - oy;
		const uint8_t sc = (ls >= 0 && ls < R) ? surface_colour(&P, (uint32_t)(ox + lx), (uint32_t)(oz + lz), h) : 0;
		any |= top > 0;
		uint8_t *p = chunk + ((size_t)lz << (2 * rb)) + lx;
		for (int ly = 0; ly < R; ly++) p[(size_t)ly << rb] = ly == ls ? sc : (ly < top ? 21 : 0);
	}
	// trees whose canopy (reach +-4) or trunk can touch this chunk: lattice cells in the host's order
	int32_t g0x = (ox - 16) / 10 - 1, g1x = (ox + R + 16) / 10 + 1, g0z = (oz - 16) / 10 - 1, g1z = (oz + R + 16) / 10 + 1;
	if (g0x < 0) g0x = 0;
	if (g0z < 0) g0z = 0;
	const int nxc = g1x - g0x + 1, ncell = nxc * (g1z - g0z + 1);
	for (int c = tid; c < ncell && c < kMaxCells; c += kGenThreads) {
		TreeRec t; t.ok = 0;
		int32_t tx, tz, ty;
		if (tree_at(&P, g0x + c % nxc, g0z + c / nxc, &tx, &tz, &ty) && !(ty + 16 < oy || ty >= oy + R)) { t.x = tx; t.y = ty; t.z = tz; t.ok = 1; }
		trees[c] = t;
	}
	__syncthreads();
	if (tid < 32) {
		int wrote = 0;
		for (int c = 0; c < ncell && c < kMaxCells; c++) {
			const TreeRec t = trees[c];
			if (!t.ok) continue;
			// canopy first, then trunk + top leaves (the trunk overwrites the canopy cells it passes through)
			for (int phase = 0; phase < 2; phase++) {
				const int kbeg = phase ? VPW_TREE_CANOPY : 0, kend = phase ? VPW_TREE_VOXELS : VPW_TREE_CANOPY;
				for (int k0 = kbeg; k0 < kend; k0 += 32) {
					const int k = k0 + tid;
					int32_t x, y, z; uint8_t v;
					if (k < kend && tree_voxel(&P, t.x, t.y, t.z, k, &x, &y, &z, &v)) {
						const int32_t lx = x - ox, ly = y - oy, lz = z - oz;
						if (lx >= 0 && ly >= 0 && lz >= 0 && lx < R && ly < R && lz < R && y >= 2) {          // edit.c:151
							chunk[(((size_t)lz << rb | (size_t)ly) << rb) | (size_t)lx] = v;
							wrote = 1;
						}
					}
					__syncwarp();
				}
			}
		}
		if (wrote) s_written = 1;
	}
	const int solid_any = __syncthreads_or(any);
	if (tid == 0) solid[blockIdx.x] = (solid_any || s_written) ? 1u : 0u;
}

// staging chunk k -> pool slot slots[k] (negative: skipped)
__global__ void k_scatter_chunks(const uint8_t *__restrict__ staging, const int32_t *__restrict__ slots, uint8_t *__restrict__ pool, uint32_t n16)
{
	const int32_t s = slots[blockIdx.x];
	if (s < 0) return;
	const uint4 *src = reinterpret_cast<const uint4 *>(staging) + (size_t)blockIdx.x * n16;
	uint4 *dst = reinterpret_cast<uint4 *>(pool) + (size_t)s * n16;
	for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
}

// One z row of the height map from resident voxels.  table[(row * ny + cy) * nx + cx] = chunk pointer or null for the
// chunk rows [row0, row0 + nrows); blockIdx.x = z - z_first.  blockDim.x = Y (<= 1024).
template <int RB>
__global__ void __launch_bounds__(1024)
k_shadow_rows(const uint8_t *const *__restrict__ table, int nx, int ny, uint32_t row0, uint32_t z_first, uint16_t *__restrict__ rows_out)
{
	constexpr int R = 1 << RB;
	extern __shared__ uint16_t row[];
	const uint32_t X = (uint32_t)nx << RB, Y = (uint32_t)ny << RB, SH = X + Y;
	const uint32_t z = z_first + blockIdx.x;
	const uint32_t y = threadIdx.x, cy = y >> RB, ly = y & (R - 1), lz = z & (R - 1);
	for (uint32_t i = threadIdx.x; i < SH + 2; i += blockDim.x) row[i] = 0;
	__syncthreads();
	const uint8_t *const *trow = table + ((size_t)((z >> RB) - row0) * ny + cy) * nx;
	const uint32_t lim = y + 1;
	for (int cx = 0; cx < nx; cx++) {
		const uint8_t *c = trow[cx];
		uint4 q[R / 16];
		#pragma unroll
		for (int k = 0; k < R / 16; k++) q[k] = c ? __ldg(reinterpret_cast<const uint4 *>(c + (((size_t)lz << RB | ly) << RB)) + k) : make_uint4(0, 0, 0, 0);
		const uint32_t base = ((uint32_t)cx << RB) + y;
		#pragma unroll
		for (int k = 0; k < R / 16; k++) {
			const uint32_t wd[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
			#pragma unroll
			for (int j = 0; j < 16; j++) {
				const uint32_t v = (wd[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
				const uint32_t idx = base + k * 16 + j;
				bool wr = false;
				if (v) wr = !(row[idx] >= lim || row[idx + 1] >= lim);          // shadow.h:82-86, old entries only
				__syncthreads();
				if (wr) row[idx] = (uint16_t)y;
				__syncthreads();
			}
		}
	}
	uint16_t *dst = rows_out + (size_t)blockIdx.x * SH;
	for (uint32_t i = threadIdx.x; i < SH; i += blockDim.x) dst[i] = row[i];
}

} // namespace

cudaError_t vp_launch_gen_chunks(uint32_t seed, int rb, const int bits[3], const uint32_t *d_ids, uint32_t n, uint8_t *d_out, uint32_t *d_solid, cudaStream_t s)
{
	if (!n) return cudaSuccess;
	vpw_params P;
	P.seed = seed; P.root_bitw = rb; P.bits[0] = bits[0]; P.bits[1] = bits[1]; P.bits[2] = bits[2];
	k_gen_chunks<<<n, kGenThreads, 0, s>>>(P, d_ids, d_out, d_solid);
	return cudaGetLastError();
}

cudaError_t vp_launch_scatter_chunks(int rb, const uint8_t *staging, const int32_t *d_slots, uint32_t n, uint8_t *pool, cudaStream_t s)
{
	if (!n) return cudaSuccess;
	k_scatter_chunks<<<n, 256, 0, s>>>(staging, d_slots, pool, (1u << (3 * rb)) / 16);
	return cudaGetLastError();
}

template <int RB>
static cudaError_t launch_shadow(const uint8_t *const *table, int nx, int ny, uint32_t row0, uint32_t z0, uint32_t z1, uint16_t *rows_out, cudaStream_t s)
{
	const uint32_t Y = (uint32_t)ny << RB, SH = ((uint32_t)(nx + ny) << RB);
	const size_t smem = (size_t)(SH + 2) * 2;
	cudaError_t e = cudaFuncSetAttribute(k_shadow_rows<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	k_shadow_rows<RB><<<z1 - z0, Y, smem, s>>>(table, nx, ny, row0, z0, rows_out);
	return cudaGetLastError();
}

// rows [z0,z1) from the chunk rows [row0, ...) described by `table`; needs Y <= 1024 and (X+Y+2)*2 bytes of shared memory
cudaError_t vp_launch_shadow_rows(int rb, const uint8_t *const *table, int nx, int ny, uint32_t row0, uint32_t z0, uint32_t z1, uint16_t *rows_out, cudaStream_t s)
{
	if (z1 <= z0) return cudaSuccess;
	if (((uint32_t)ny << rb) > 1024u || (size_t)(((uint32_t)(nx + ny) << rb) + 2) * 2 > 200u * 1024u) return cudaErrorInvalidConfiguration;
	switch (rb) {
	case 4: return launch_shadow<4>(table, nx, ny, row0, z0, z1, rows_out, s);
	case 5: return launch_shadow<5>(table, nx, ny, row0, z0, z1, rows_out, s);
	case 6: return launch_shadow<6>(table, nx, ny, row0, z0, z1, rows_out, s);
	case 7: return launch_shadow<7>(table, nx, ny, row0, z0, z1, rows_out, s);
	default: return cudaErrorInvalidValue;
	}
}
