// vp_nodes.cu -- LOD-node aggregation of the splat lists (the step right after the path).
//
// Replaces the gather loops of gfx_update_svl (gfx/vsplat.c:209-323): for LOD level `lod` the world is cut into
// octree nodes of 2^lod chunks per axis (node = chunk offset >> lod, :216-222; node index = flatten3 with bit
// widths max_bitw - min(lod, max_bitw), :214,229); a node's vertex buffer is the concatenation of the level-`lod`
// segment of every member chunk's splat list, members visited x outer, y, z inner (:264-266, :286-288), each
// segment starting after the chunk's lower levels (:297-300).  The reference rebuilds a node on the CPU and
// re-uploads it whenever one member changes; here all nodes of a level are gathered on the device in one
// launch straight from the splat arena the rebuild wrote.
//
// Plan kernel, one CTA per node: member sizes -> block prefix scan (stable member order) -> one arena reservation
// -> destination of every member chunk.  Copy kernel, one CTA per chunk: coalesced 8-byte copies (every splat is
// 8 bytes, so all segment boundaries are 8-byte aligned).
#include "vp_device.cuh"
using namespace vp;

namespace {

constexpr int kT = 256;

__global__ void __launch_bounds__(kT)
k_lod_nodes_plan(int lod, int bx, int by, int bz, const VpResultDev *__restrict__ chunk_res,
            VpArenaDev *__restrict__ st, VpNodeDev *__restrict__ nodes, unsigned long long *__restrict__ chunk_dst)
{
	__shared__ uint32_t s_pre[4096 + 1];          // member prefix (items); 8^4 members at most
	__shared__ uint32_t s_wsum[kT / 32];
	__shared__ unsigned long long s_off;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int obx = bx - min(lod, bx), oby = by - min(lod, by), obz = bz - min(lod, bz);
	const uint32_t node = blockIdx.x;
	const uint32_t ox = node & ((1u << obx) - 1), oy = (node >> obx) & ((1u << oby) - 1), oz = node >> (obx + oby);
	const uint32_t x0 = ox << lod, y0 = oy << lod, z0 = oz << lod;
	const uint32_t nx = min((ox + 1) << lod, 1u << bx) - x0, ny = min((oy + 1) << lod, 1u << by) - y0, nz = min((oz + 1) << lod, 1u << bz) - z0;
	const uint32_t nm = nx * ny * nz;             // members, order: x outer, y, z inner
	auto member_chunk = [&](uint32_t m) -> uint32_t {
		const uint32_t dz = m % nz, dy = (m / nz) % ny, dx = m / (nz * ny);
		return ((((z0 + dz) << by) | (y0 + dy)) << bx) | (x0 + dx);
	};
	// member sizes and their exclusive prefix
	constexpr int IPT = 4096 / kT;
	uint32_t v[IPT], sum = 0;
	#pragma unroll
	for (int k = 0; k < IPT; k++) {
		const uint32_t m = tid * IPT + k;
		v[k] = m < nm ? chunk_res[member_chunk(m)].svl_items[lod] : 0u;
		sum += v[k];
	}
	uint32_t inc = sum;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
	if (lane == 31) s_wsum[warp] = inc;
	__syncthreads();
	uint32_t pre = inc - sum;
	for (int k = 0; k < warp; k++) pre += s_wsum[k];
	#pragma unroll
	for (int k = 0; k < IPT; k++) { const uint32_t m = tid * IPT + k; if (m <= nm) s_pre[m] = pre; pre += v[k]; }
	if (tid == kT - 1) s_pre[nm] = pre;
	__syncthreads();
	const uint32_t total = s_pre[nm];             // int16 items
	if (tid == 0) {
		unsigned long long off = 0;
		const unsigned long long bytes = (unsigned long long)total * 2ull;
		if (total) {
			off = atomicAdd(&st->cursor, bytes);
			if (off + bytes > st->capacity) { atomicExch(&st->overflow, 1u); off = ~0ull; }
		}
		s_off = off;
		nodes[node].offset = off;
		nodes[node].items = total;                // GeometrySVL.vbo_items (vsplat.c:325)
		nodes[node].members = nm;
	}
	__syncthreads();
	// plan only: every member learns where its segment goes; the copy runs with one CTA per CHUNK so that the
	// few, huge nodes of the high levels are gathered by the whole GPU instead of one CTA each
	for (uint32_t m = tid; m < nm; m += kT)
		chunk_dst[member_chunk(m)] = (total && s_off != ~0ull && s_pre[m + 1] != s_pre[m]) ? s_off + (unsigned long long)s_pre[m] * 2ull : ~0ull;
}

__global__ void __launch_bounds__(kT)
k_lod_nodes_copy(int lod, const VpResultDev *__restrict__ chunk_res, const uint8_t *__restrict__ splat_arena,
                 uint8_t *__restrict__ node_arena, const unsigned long long *__restrict__ chunk_dst)
{
	const unsigned long long off = chunk_dst[blockIdx.x];
	if (off == ~0ull) return;
	const VpResultDev &r = chunk_res[blockIdx.x];
	uint32_t start = 0;
	for (int l = 0; l < lod; l++) start += r.svl_items[l];           // vsplat.c:297-300
	const uint32_t n8 = r.svl_items[lod] >> 2;                       // 8-byte splat records
	const unsigned long long *src = reinterpret_cast<const unsigned long long *>(splat_arena + r.svl_offset) + (start >> 2);
	unsigned long long *dst = reinterpret_cast<unsigned long long *>(node_arena + off);
	for (uint32_t i = threadIdx.x; i < n8; i += kT) dst[i] = __ldg(src + i);
}

} // namespace

cudaError_t vp_launch_lod_nodes(int lod, const int bits[3], uint32_t n_nodes, const VpResultDev *d_chunk_res, const uint8_t *d_splat_arena,
                                uint8_t *d_node_arena, VpArenaDev *state, VpNodeDev *d_nodes, unsigned long long *d_chunk_dst, cudaStream_t s)
{
	if (!n_nodes) return cudaSuccess;
	k_lod_nodes_plan<<<n_nodes, kT, 0, s>>>(lod, bits[0], bits[1], bits[2], d_chunk_res, state, d_nodes, d_chunk_dst);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	k_lod_nodes_copy<<<1u << (bits[0] + bits[1] + bits[2]), kT, 0, s>>>(lod, d_chunk_res, d_splat_arena, d_node_arena, d_chunk_dst);
	return cudaGetLastError();
}
