// vp_mesh.cu -- near-field quad mesh (VBO + IBO) for a batch of chunks: two kernels, no communication between CTAs.
//
// Replaces, byte for byte, the mesh branch of the reference dispatcher (chunkset.c:339-343):
//   chunk_make_mesh (mesher.c:184-357) with sample_ao / sample_ao_border (mesher.c:72-171).
//
// k_mesh_count -- one CTA per 16-slice z-slab of a chunk:
//   1. streams its slab plus the slices below and above (own chunk or the -z / +z neighbour) through a TMA ring and
//      packs them to occupancy rows; the y = -1 / y = R rows come from the +-y (and diagonal) neighbours via small TMA
//      copies, the x = -1 / x = R bits from the neighbours' x-face planes -- so the tile is the full (R+2)^3
//      neighbourhood the AO samples need, with air outside the world (edit.c:22-25);
//   2. a face exists between voxel A and A+e_i when exactly one of them is solid (mesher.c:236): three XORs of bit
//      rows.  Faces are ordered by A (z,y,x) and then by axis, so the three face rows of a voxel row are
//      bit-interleaved (index 3x+i) 16 voxels at a time through a 256-entry spread table; units of 48 interleaved bits
//      are counted with popc, groups of 32 units with REDUX, one warp scans the group counts;
//   3. the occupancy tile (19 KB for 64^3 chunks, a quarter of the voxels it was made from) and the group prefixes go
//      to a scratch; the LAST slab of a chunk to finish (arrival counter) reserves the chunk's [VBO | IBO] buffer with
//      one atomicAdd and writes every slab's first face.  Nobody waits.
// k_mesh_emit -- one CTA per slab with faces: fetches the tile with bulk copies and emits a few slices at a time:
//   the non-empty units are compacted into entries {first slot, unit, face bits}, a table gives the first entry of
//   every round of 32 consecutive slots, slot -> entry by popc over the mask of entry starts (the machinery of
//   vp_splat.cu); per face: AO from 8 occupancy bits, shadow / diamond bits from the height map, colour byte from L2,
//   4 vertices (32 B) + 6 indices (24 B) stored contiguously by rank.
// (The splat kernel stages its 8-byte records and lets the last slab move them; here a face is 56 bytes of output for a
// few bits of input, so carrying the bit tile to a second kernel is the cheaper way to learn the chunk totals first.)
#include "vp_device.cuh"
#include <cstddef>
using namespace vp;

namespace {

constexpr int kRing = 4;
constexpr int kConsumerWarps = 4;
constexpr int kThreads = 256;          // count kernel: 4 consumer warps, 3 helper warps (halo rows / x bits), 1 producer warp
constexpr int kWarps = kThreads / 32;
constexpr int kProducerWarp = kWarps - 1;
constexpr int kEmitThreads = 256;
constexpr int kEmitWarps = kEmitThreads / 32;
constexpr int kMeshRec = 8;            // uint32 per slab record: [0] faces, [1] first face inside the chunk, [2] faces of the chunk, [4,5] chunk offset in the arena

template <int RB> struct MGeo {
	static constexpr int R = 1 << RB;
	static constexpr int ZS = 16;
	static constexpr int CL = R / ZS;
	static constexpr int NW = R > 64 ? R / 64 : 1;       // main 64-bit words per occupancy row
	static constexpr int RW = NW + 1;                    // + 1 word holding the x = -1 (bit 0) and x = R (bit 1) cells
	static constexpr int SLICE = R * R;
	static constexpr int TILE = SLICE < 4096 ? SLICE : 4096;
	static constexpr int TPS = SLICE / TILE;
	static constexpr int RPT = TILE / R;
	static constexpr int NSL = ZS + 2;                   // slices z0-1 .. z0+ZS
	static constexpr int NT = NSL * TPS;
	static constexpr int LPR = R / 16;
	static constexpr int ROWS = R + 2;                   // y = -1 .. R
	static constexpr int UPR = R / 16;                   // units (16 voxels -> 48 interleaved face bits) per voxel row
	static constexpr int NU = ZS * R * UPR;              // units per CTA
	static constexpr int NG = NU / 32;                   // groups of 32 units
	static constexpr int OCC_WORDS = NSL * ROWS * RW;
	static constexpr int OCC_STRIDE = (OCC_WORDS + 1) / 2 * 2;          // uint64 words, 16-byte multiple for bulk copies
	// group record: [NG + 1] uint32 exclusive prefix of the faces, then [NG + 1] uint16 exclusive prefix of the non-empty units
	static constexpr int GP_STRIDE = ((NG + 1) + (NG + 2) / 2 + 3) / 4 * 4;     // uint32 words
	// emission passes: SUBZ slices (at most 1024 units) at a time
	static constexpr int SUBZ = (1024 / (R * UPR)) < 1 ? 1 : ((1024 / (R * UPR)) > ZS ? ZS : (1024 / (R * UPR)));
	static constexpr int PU = SUBZ * R * UPR;            // units per pass
	static constexpr int PG = PU / 32;                   // groups per pass
	static constexpr int NPASS = ZS / SUBZ;
	static constexpr int TBL_N = PU * 48 / 32 + 2;       // rounds of a pass: a unit holds at most 48 faces
	// count kernel shared memory (bytes)
	static constexpr int RING_BYTES = kRing * TILE;
	static constexpr int OFF_YH = (RING_BYTES + 127) / 128 * 128;      // y-halo staging: [2][NSL][R] bytes
	static constexpr int YH_BYTES = 2 * NSL * R;
	static constexpr int OFF_OCC = (OFF_YH + YH_BYTES + 127) / 128 * 128;
	static constexpr int OFF_GP = OFF_OCC + OCC_STRIDE * 8;
	static constexpr int OFF_LUT = OFF_GP + GP_STRIDE * 4;
	static constexpr int OFF_BARS = OFF_LUT + 256 * 4;
	static constexpr int OFF_MISC = OFF_BARS + (2 * kRing + 2) * 8;
	static constexpr int SMEM = OFF_MISC + 128;
	// emit kernel shared memory (bytes)
	static constexpr int E_OFF_OCC = 0;
	static constexpr int E_OFF_GP = OCC_STRIDE * 8;
	static constexpr int E_OFF_LUT = E_OFF_GP + GP_STRIDE * 4;
	static constexpr int E_OFF_ENT = E_OFF_LUT + 256 * 4;               // uint4 per non-empty unit of the pass
	static constexpr int E_OFF_TBL = E_OFF_ENT + PU * 16;
	static constexpr int E_OFF_MISC = E_OFF_TBL + (TBL_N * 2 + 15) / 16 * 16;
	static constexpr int E_SMEM = E_OFF_MISC + 64;
	static_assert(NU <= 65535, "unit prefixes are stored in 16 bits");
	static_assert(PU <= 65535, "entry indices are stored in 16 bits");
};

struct MMisc {
	int32_t slot[27];               // neighbourhood slots [dz+1][dy+1][dx+1]
	uint32_t last;
};
static_assert(sizeof(MMisc) <= 128, "MMisc must fit the reserved 128 bytes");

struct EMisc { uint64_t bar; int32_t slot[4]; };      // emit kernel: own, +x, +y, +z slots

// quad corner offsets per axis, 3 bits per vertex (bit0 = x, bit1 = y, bit2 = z): mesher.c:19-33
__device__ __constant__ uint8_t c_corner[3][4] = { {1, 5, 7, 3}, {2, 3, 7, 6}, {5, 4, 6, 7} };
// index patterns [normal][rotated][6]: mesher.c:53-65
__device__ __constant__ uint8_t c_index[2][2][6] = { { {0, 3, 1, 2, 1, 3}, {3, 2, 0, 1, 0, 2} }, { {1, 3, 0, 3, 1, 2}, {0, 2, 3, 2, 0, 1} } };

template <int RB> struct MCtx {
	using G = MGeo<RB>;
	const VpWorldDev &w;
	const uint64_t *occ;
	const int32_t *slot4;           // own, +x, +y, +z slots (emission only)
	int z0;
	uint32_t ox, oy, oz;

	// occupancy of local cell (x,y,z) with x,y in [-1,R], z in [z0-1, z0+ZS]
	__device__ __forceinline__ uint32_t occ_at(int x, int y, int z) const
	{
		const uint64_t *row = occ + (size_t)((z - z0 + 1) * G::ROWS + (y + 1)) * G::RW;
		if (x < 0) return (uint32_t)row[G::NW] & 1u;
		if (x >= G::R) return ((uint32_t)row[G::NW] >> 1) & 1u;
		return (uint32_t)(row[x >> 6] >> (x & 63)) & 1u;
	}

	// 16 occupancy bits x0 .. x0+15 of row (y,z) and the bit of x0+16 (for the +x face test)
	__device__ __forceinline__ uint32_t occ17(int x0, int y, int z) const
	{
		const uint64_t *row = occ + (size_t)((z - z0 + 1) * G::ROWS + (y + 1)) * G::RW;
		const uint64_t wd = row[x0 >> 6];
		uint32_t v = (uint32_t)(wd >> (x0 & 63)) & 0xFFFFu;
		uint32_t nxt;
		if (x0 + 16 >= G::R) nxt = ((uint32_t)row[G::NW] >> 1) & 1u;
		else if (((x0 + 16) & 63) == 0) nxt = (uint32_t)row[(x0 + 16) >> 6] & 1u;
		else nxt = (uint32_t)(wd >> ((x0 & 63) + 16)) & 1u;
		return v | (nxt << 16);
	}

	// interleaved face bits (index 3*dx + axis) of unit u = ((zi * R + y) * UPR + j): 16 voxels x0 = 16 j ..
	__device__ __forceinline__ uint64_t unit_faces(const uint32_t *lut, int u) const
	{
		const int j = u % G::UPR, r = u / G::UPR, y = r & (G::R - 1), z = z0 + (r >> RB), x0 = j * 16;
		const uint32_t a = occ17(x0, y, z);
		const uint32_t o = a & 0xFFFFu;
		const uint32_t fx = o ^ (a >> 1);
		const uint32_t fy = o ^ (occ17(x0, y + 1, z) & 0xFFFFu);
		const uint32_t fz = o ^ (occ17(x0, y, z + 1) & 0xFFFFu);
		if ((fx | fy | fz) == 0u) return 0ull;
		const uint64_t sx = (uint64_t)lut[fx & 255u] | ((uint64_t)lut[fx >> 8] << 24);
		const uint64_t sy = (uint64_t)lut[fy & 255u] | ((uint64_t)lut[fy >> 8] << 24);
		const uint64_t sz = (uint64_t)lut[fz & 255u] | ((uint64_t)lut[fz >> 8] << 24);
		return sx | (sy << 1) | (sz << 2);
	}

	// voxel byte of local cell (x,y,z), exactly one coordinate may be R (the +neighbour's first cell)
	__device__ __forceinline__ uint32_t voxel(int x, int y, int z) const
	{
		constexpr int R = G::R;
		const size_t N = (size_t)R * R * R;
		if (x >= R) return __ldg(w.xlo_pool + (size_t)slot4[1] * R * R + (size_t)z * R + y);
		if (y >= R) return __ldg(w.vox_pool + (size_t)slot4[2] * N + (size_t)z * R * R + x);
		if (z >= R) return __ldg(w.vox_pool + (size_t)slot4[3] * N + (size_t)y * R + x);
		return __ldg(w.vox_pool + (size_t)slot4[0] * N + ((size_t)z * R + y) * R + x);
	}

	// One quad: face between A = (x,y,z) and A + e_i, rank = index of the face inside the chunk.
	// (Scalar component arithmetic instead of arrays indexed by the axis keeps everything in registers.)
	__device__ __forceinline__ void emit_face(uint4 *vbo, uint2 *ibo, uint32_t rank, int x, int y, int z, int i) const
	{
		const int ex = i == 0, ey = i == 1, ez = i == 2;                                   // e_i
		const uint32_t a_occ = occ_at(x, y, z);
		const int normal = a_occ ? 0 : 1;                                                  // mesher.c:292
		// AIR = A + e_i if B is air (A solid) else A; BLOCK the other one (mesher.c:239-244)
		const int ax = x + (a_occ ? ex : 0), ay = y + (a_occ ? ey : 0), az = z + (a_occ ? ez : 0);
		const int bx = x + (a_occ ? 0 : ex), by = y + (a_occ ? 0 : ey), bz = z + (a_occ ? 0 : ez);
		const uint32_t colour = voxel(bx, by, bz);                                         // A|B (mesher.c:333)
		// the two in-plane axes u0 < u1 (mesher.c:84-98): i=0 -> (y,z), i=1 -> (x,z), i=2 -> (x,y)
		const int u0x = !ex, u0y = ex, u1y = ez, u1z = !ez;
		#define VP_OCC(d0, d1) occ_at(ax + (d0) * u0x, ay + (d0) * u0y + (d1) * u1y, az + (d1) * u1z)
		const uint32_t n0 = VP_OCC(-1, 0), n1 = VP_OCC(0, -1), n2 = VP_OCC(1, 0), n3 = VP_OCC(0, 1);
		const uint32_t c0 = VP_OCC(-1, -1), c1 = VP_OCC(1, -1), c2 = VP_OCC(1, 1), c3 = VP_OCC(-1, 1);
		#undef VP_OCC
		const uint32_t q0 = (n0 + n1) | c0, q1 = (n1 + n2) | c1, q2 = (n2 + n3) | c2, q3 = (n3 + n0) | c3;
		// axis-dependent assignment of the four AO terms to the quad vertices (mesher.c:257-272)
		uint32_t v0, v1, v2, v3;
		if (i == 0) { v0 = q0; v3 = q1; v2 = q2; v1 = q3; }
		else if (i == 1) { v0 = q0; v1 = q1; v2 = q2; v3 = q3; }
		else { v1 = q0; v0 = q1; v3 = q2; v2 = q3; }
		const int rotated = (v0 + v2 < v1 + v3) ? 1 : 0;                                   // mesher.c:274-276
		uint32_t wx = ox + (uint32_t)bx, wy = oy + (uint32_t)by, wz = oz + (uint32_t)bz;
		uint32_t shadow = 0, diamond = 3;
		if (i == 0 && normal) shadow = (uint32_t)shadow_pair(w, wx, wy, wz, -1);           // mesher.c:295-298
		else if (i == 1 && !normal) shadow = (uint32_t)shadow_pair(w, wx, wy, wz, 1);
		if (normal) { wx -= (uint32_t)ex; wy -= (uint32_t)ey; wz -= (uint32_t)ez; }        // mesher.c:300
		if (i == 2) {                                                                      // mesher.c:306-317
			const uint32_t sx = ox + (uint32_t)ax, sy = oy + (uint32_t)ay, sz = oz + (uint32_t)az;
			shadow = 0;
			diamond = ((uint32_t)!shadow_pair(w, sx, sy, sz, 1)) << 1;
			diamond |= (uint32_t)!shadow_pair(w, sx, sy - 1u, sz, 1);
		}
		const uint32_t common = colour | ((uint32_t)(i + 3 * normal) << 8) | (shadow << 13) | (diamond << 14);
		const uint32_t vao[4] = {v0, v1, v2, v3};
		uint32_t lo[4], hi[4];
		#pragma unroll
		for (int t = 0; t < 4; t++) {
			const uint32_t cr = c_corner[i][t];
			const uint32_t vx = (wx + (cr & 1u)) & 0xFFFFu, vy = (wy + ((cr >> 1) & 1u)) & 0xFFFFu, vz = (wz + ((cr >> 2) & 1u)) & 0xFFFFu;
			const uint32_t vd = (common | (vao[t] << 6) | ((uint32_t)t << 11)) & 0xFFFFu;  // mesher.c:332-338
			lo[t] = vx | (vy << 16); hi[t] = vz | (vd << 16);
		}
		vbo[0] = make_uint4(lo[0], hi[0], lo[1], hi[1]);
		vbo[1] = make_uint4(lo[2], hi[2], lo[3], hi[3]);
		const uint32_t b4 = rank * 4u;                                                     // mesher.c:345-349
		const uint8_t *ix = c_index[normal][rotated];
		ibo[0] = make_uint2(b4 + ix[0], b4 + ix[1]);
		ibo[1] = make_uint2(b4 + ix[2], b4 + ix[3]);
		ibo[2] = make_uint2(b4 + ix[4], b4 + ix[5]);
	}
};

__device__ __forceinline__ int select64m(uint32_t lo, uint32_t hi, uint32_t cl, uint32_t k)
{
	uint32_t v = lo, c; int pos = 0;
	if (k >= cl) { k -= cl; v = hi; pos = 32; }
	c = __popc(v & 0xFFFFu); if (k >= c) { k -= c; v >>= 16; pos += 16; }
	c = __popc(v & 0xFFu);   if (k >= c) { k -= c; v >>= 8;  pos += 8; }
	c = __popc(v & 0xFu);    if (k >= c) { k -= c; v >>= 4;  pos += 4; }
	c = __popc(v & 0x3u);    if (k >= c) { k -= c; v >>= 2;  pos += 2; }
	c = v & 1u;              if (k >= c) { pos += 1; }
	return pos;
}

// spread an 8-bit value to every third bit
__device__ __forceinline__ void fill_spread_lut(uint32_t *lut, int tid, int nthreads)
{
	for (int i = tid; i < 256; i += nthreads) {
		uint32_t v = 0;
		#pragma unroll
		for (int b = 0; b < 8; b++) v |= ((uint32_t)(i >> b) & 1u) << (3 * b);
		lut[i] = v;
	}
}

// Scratch between the two kernels (device pointers, sized by vp_mesh_scratch_bytes).
struct MeshScratch {
	uint32_t *arrived;              // [cap chunks]         slabs of the chunk that finished counting (self-resetting)
	uint64_t *occ;                  // [slabs][OCC_STRIDE]  occupancy tile of every slab with faces
	uint32_t *gp;                   // [slabs][GP_STRIDE]   packed exclusive prefix of the per-group face / unit counts
	uint32_t *rec;                  // [slabs][kMeshRec]
};

inline size_t arrived_region_bytes(uint32_t cap_chunks) { return ((size_t)cap_chunks * 4 + 255) / 256 * 256; }

template <int RB>
__host__ __device__ __forceinline__ MeshScratch carve_scratch(uint8_t *base, size_t arrived_bytes, uint32_t cap_chunks)
{
	using G = MGeo<RB>;
	const size_t slabs = (size_t)cap_chunks * G::CL;
	MeshScratch sc;
	sc.arrived = reinterpret_cast<uint32_t *>(base);
	base += arrived_bytes;
	sc.occ = reinterpret_cast<uint64_t *>(base);
	sc.gp = reinterpret_cast<uint32_t *>(base + slabs * G::OCC_STRIDE * 8);
	sc.rec = sc.gp + slabs * G::GP_STRIDE;
	return sc;
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel 1: stream the (R+2)^2 x 18 neighbourhood of a slab, pack it to bits, count the faces.
// ------------------------------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(kThreads, RB == 7 ? 2 : 5)
k_mesh_count(const VpWorldDev w, const uint32_t *__restrict__ ids, uint8_t *__restrict__ scratch, size_t arrived_bytes, uint32_t cap_chunks,
             VpResultDev *__restrict__ results, const uint32_t *__restrict__ result_pos, VpArenaDev *__restrict__ st)
{
	using G = MGeo<RB>;
	constexpr int R = G::R, ZS = G::ZS, CL = G::CL, NW = G::NW, RW = G::RW, TILE = G::TILE, TPS = G::TPS, NT = G::NT, NSL = G::NSL;
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t *ring = smem;
	uint8_t *yh = smem + G::OFF_YH;
	uint64_t *occ = reinterpret_cast<uint64_t *>(smem + G::OFF_OCC);
	uint32_t *gpre = reinterpret_cast<uint32_t *>(smem + G::OFF_GP);
	uint32_t *lut = reinterpret_cast<uint32_t *>(smem + G::OFF_LUT);
	uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + G::OFF_BARS);
	uint64_t *bar_empty = bar_full + kRing;
	uint64_t *bar_halo = bar_empty + kRing;
	MMisc *misc = reinterpret_cast<MMisc *>(smem + G::OFF_MISC);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int crank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
	const uint32_t chunk_i = blockIdx.x / CL;
	const uint32_t cid = ids[chunk_i];
	const int cx = (int)(cid & ((1u << w.bits[0]) - 1)), cy = (int)((cid >> w.bits[0]) & ((1u << w.bits[1]) - 1));
	const int cz = (int)(cid >> (w.bits[0] + w.bits[1]));
	const int z0 = crank * ZS;
	const size_t N = (size_t)R * R * R;
	const MeshScratch sc = carve_scratch<RB>(scratch, arrived_bytes, cap_chunks);
	uint32_t *rec = sc.rec + (size_t)blockIdx.x * kMeshRec;

	// ---- phase 0: neighbourhood slots, barriers, spread table, zeroed occupancy ----------------------
	if (tid < 27) misc->slot[tid] = chunk_slot(w, cx + tid % 3 - 1, cy + (tid / 3) % 3 - 1, cz + tid / 9 - 1);
	if (tid == kProducerWarp * 32) {
		for (int i = 0; i < kRing; i++) { mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, kConsumerWarps / kRing); }
		mbar_init(bar_halo, 1);
		mbar_fence_init();
	}
	fill_spread_lut(lut, tid, kThreads);
	for (int i = tid; i < G::OCC_STRIDE; i += kThreads) occ[i] = 0;
	__syncthreads();
	auto slot_of = [&](int dx, int dy, int dz) -> int { return misc->slot[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)]; };
	// chunk-row offset and wrapped local z of streamed slice s (z = z0 - 1 + s)
	auto slice_dz = [&](int s) -> int { const int z = z0 - 1 + s; return z < 0 ? -1 : (z >= R ? 1 : 0); };
	auto slice_lz = [&](int s) -> int { return (z0 - 1 + s) & (R - 1); };
	auto slice_src = [&](int s) -> const uint8_t * {
		const int sl = slot_of(0, 0, slice_dz(s));
		return sl >= 0 ? w.vox_pool + (size_t)sl * N + (size_t)slice_lz(s) * R * R : nullptr;
	};
	// y-halo rows: side 0 = y = -1 (row R-1 of the -y neighbour), side 1 = y = R (row 0 of the +y neighbour)
	auto yrow_src = [&](int side, int s) -> const uint8_t * {
		const int sl = slot_of(0, side ? 1 : -1, slice_dz(s));
		return sl >= 0 ? w.vox_pool + (size_t)sl * N + ((size_t)slice_lz(s) * R + (side ? 0 : R - 1)) * R : nullptr;
	};

	// ---- phase 1: TMA producer / byte->bit consumers / helpers -------------------------------------------------
	if (warp == kProducerWarp) {
		if (lane == 0) {
			uint32_t hb = 0;
			for (int k = 0; k < 2 * NSL; k++) if (yrow_src(k / NSL, k % NSL)) hb += R;
			if (hb) {
				mbar_arrive_expect_tx(bar_halo, hb);
				for (int k = 0; k < 2 * NSL; k++) {
					const uint8_t *src = yrow_src(k / NSL, k % NSL);
					if (src) tma_load_1d(yh + k * R, src, R, bar_halo);
				}
			} else {
				mbar_arrive(bar_halo);
			}
			for (int t = 0; t < NT; t++) {
				const int b = t % kRing, u = t / kRing;
				if (u > 0) mbar_wait(bar_empty + b, (u - 1) & 1);
				const uint8_t *src = slice_src(t / TPS);
				if (src) {
					mbar_arrive_expect_tx(bar_full + b, TILE);
					tma_load_1d(ring + b * TILE, src + (size_t)(t % TPS) * TILE, TILE, bar_full + b);
				} else {
					mbar_arrive(bar_full + b);
				}
			}
		}
	} else if (warp < kConsumerWarps) {
		constexpr int WPS = kConsumerWarps / kRing, PART = TILE / WPS;
		const int b = warp % kRing, hpart = warp / kRing;
		for (int t = b; t < NT; t += kRing) {
			const int u = t / kRing, s = t / TPS, part = t % TPS;
			mbar_wait(bar_full + b, u & 1);
			if (slice_src(s)) {
				const uint8_t *tb = ring + b * TILE;
				uint64_t *orow = occ + (size_t)(s * G::ROWS + 1 + part * G::RPT) * RW;      // +1: row index y+1
				#pragma unroll 4
				for (int off = hpart * PART + lane * 16; off < (hpart + 1) * PART; off += 512) {
					const uint4 q4 = *reinterpret_cast<const uint4 *>(tb + off);
					if (PART >= 512 && !__any_sync(0xffffffffu, (q4.x | q4.y | q4.z | q4.w) != 0u)) continue;
					const uint32_t m = nz16(q4);
					const int row = off / R, bo = off % R;
					if (R >= 32) {
						uint32_t v = m << (bo & 16);
						v |= __shfl_xor_sync(0xffffffffu, v, 1);
						if (!(lane & 1)) reinterpret_cast<uint32_t *>(orow + row * RW)[bo >> 5] = v;
					} else {
						reinterpret_cast<uint32_t *>(orow + row * RW)[0] = m;
					}
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_empty + b);
		}
	} else {
		// helper warps, while the slab streams: x = -1 / x = R cells of every row (y = -1 .. R) of every slice, from the
		// x-face planes of the 9+9 chunks on either side (two byte loads per row, contiguous in y) ...
		constexpr int kHelpers = (kWarps - 1 - kConsumerWarps) * 32;
		const int ht = tid - kConsumerWarps * 32;
		for (int f = ht; f < NSL * G::ROWS; f += kHelpers) {
			const int s = f / G::ROWS, yy = f % G::ROWS, y = yy - 1;
			const int dy = y < 0 ? -1 : (y >= R ? 1 : 0), ly = y & (R - 1), dz = slice_dz(s), lz = slice_lz(s);
			const int sm = slot_of(-1, dy, dz), sp = slot_of(1, dy, dz);
			uint32_t e = 0;
			if (sm >= 0) e |= __ldg(w.xhi_pool + (size_t)sm * R * R + (size_t)lz * R + ly) ? 1u : 0u;
			if (sp >= 0) e |= __ldg(w.xlo_pool + (size_t)sp * R * R + (size_t)lz * R + ly) ? 2u : 0u;
			occ[(size_t)f * RW + NW] = e;
		}
		// ... and the y-halo rows -> occupancy rows 0 (y = -1) and R+1 (y = R) of every slice
		mbar_wait(bar_halo, 0);
		for (int f0 = (warp - kConsumerWarps) * 32; f0 < 2 * NSL * G::LPR; f0 += kHelpers) {      // whole warps: the pair shuffle needs every lane
			const int f = f0 + lane;
			const bool valid = f < 2 * NSL * G::LPR;
			const int k = valid ? f / G::LPR : 0, bo = (f % G::LPR) * 16, side = k / NSL, s = k % NSL;
			uint32_t m = (valid && yrow_src(side, s)) ? nz16(*reinterpret_cast<const uint4 *>(yh + k * R + bo)) : 0u;
			uint64_t *dst = occ + (size_t)(s * G::ROWS + (side ? R + 1 : 0)) * RW;
			if (R >= 32) {
				uint32_t v = m << (bo & 16);
				v |= __shfl_xor_sync(0xffffffffu, v, 1);
				if (valid && !(lane & 1)) reinterpret_cast<uint32_t *>(dst)[bo >> 5] = v;
			} else if (valid) {
				reinterpret_cast<uint32_t *>(dst)[0] = m;
			}
		}
	}
	__syncthreads();

	MCtx<RB> mc{w, occ, nullptr, z0, (uint32_t)cx << RB, (uint32_t)cy << RB, (uint32_t)cz << RB};

	// ---- phase 2: face counts per group of 32 units, packed exclusive prefix over the groups --------------------
	uint16_t *gne = reinterpret_cast<uint16_t *>(gpre + G::NG + 1);
	for (int g = warp; g < G::NG; g += kWarps) {
		uint32_t c = __popcll(mc.unit_faces(lut, g * 32 + lane));
		const uint32_t ne = __popc(__ballot_sync(0xffffffffu, c != 0u));
		c = __reduce_add_sync(0xffffffffu, c);
		if (lane == 0) { gpre[g] = c; gne[g] = (uint16_t)ne; }
	}
	__syncthreads();
	if (warp == 0) {
		// one scan for both counts: faces in the low half, non-empty units in the high half of a 64-bit value
		constexpr int IPT = (G::NG + 31) / 32;
		unsigned long long v[IPT], sum = 0;
		#pragma unroll
		for (int k = 0; k < IPT; k++) { const int g = lane * IPT + k; v[k] = g < G::NG ? ((unsigned long long)gpre[g] | ((unsigned long long)gne[g] << 32)) : 0ull; sum += v[k]; }
		unsigned long long inc = sum;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
		unsigned long long pre = inc - sum;
		__syncwarp();
		#pragma unroll
		for (int k = 0; k < IPT; k++) { const int g = lane * IPT + k; if (g < G::NG) { gpre[g] = (uint32_t)pre; gne[g] = (uint16_t)(pre >> 32); } pre += v[k]; }
		if (lane == 31) { gpre[G::NG] = (uint32_t)pre; gne[G::NG] = (uint16_t)(pre >> 32); }
	}
	fence_proxy_async_smem();          // every thread: its writes to the tile become visible to the bulk-copy engine
	__syncthreads();
	const uint32_t faces = gpre[G::NG];

	// ---- phase 3: tile + prefixes to the scratch, chunk bookkeeping by the last slab ------------------------------
	if (faces != 0 && tid == 0) {
		bulk_store(sc.occ + (size_t)blockIdx.x * G::OCC_STRIDE, occ, G::OCC_STRIDE * 8);
		bulk_store(sc.gp + (size_t)blockIdx.x * G::GP_STRIDE, gpre, G::GP_STRIDE * 4);
		bulk_commit();
		bulk_wait_read();                                 // (the data itself is complete before the next kernel starts)
	}
	if (tid >= 32) return;
	VpResultDev *res = results + (result_pos ? result_pos[chunk_i] : chunk_i);
	if (lane == 0) rec[0] = faces;
	if (CL > 1) {
		uint32_t prev = 0;
		if (lane == 0) { __threadfence(); prev = atomicAdd(sc.arrived + chunk_i, 1u); }
		prev = __shfl_sync(0xffffffffu, prev, 0);
		if (prev != CL - 1) return;
		if (lane == 0) sc.arrived[chunk_i] = 0;           // ready for the next launch
		__threadfence();
	}
	// lane r < CL: faces of slab r; exclusive prefix + total by shuffles (CL <= 8)
	uint32_t *rc = sc.rec + (size_t)chunk_i * CL * kMeshRec;
	const uint32_t mine = lane < CL ? __ldcg(rc + lane * kMeshRec) : 0u;
	uint32_t inc = mine, t;
	t = __shfl_up_sync(0xffffffffu, inc, 1); if (lane >= 1) inc += t;
	t = __shfl_up_sync(0xffffffffu, inc, 2); if (lane >= 2) inc += t;
	t = __shfl_up_sync(0xffffffffu, inc, 4); if (lane >= 4) inc += t;
	const uint32_t tot = __shfl_sync(0xffffffffu, inc, CL - 1);
	unsigned long long off = 0;
	if (lane == 0 && tot) {
		const unsigned long long bytes = ((unsigned long long)tot * 56ull + 15ull) & ~15ull;      // keep every VBO 16-byte aligned
		off = atomicAdd(&st->cursor, bytes);
		if (off + bytes > st->capacity) { atomicExch(&st->overflow, 1u); off = ~0ull; }
	}
	off = __shfl_sync(0xffffffffu, off, 0);
	if (lane < CL) {
		rc[lane * kMeshRec + 1] = inc - mine;
		rc[lane * kMeshRec + 2] = tot;
		rc[lane * kMeshRec + 4] = (uint32_t)off;
		rc[lane * kMeshRec + 5] = (uint32_t)(off >> 32);
	}
	if (lane == 0) {
		res->vbo_offset = off;
		res->ibo_offset = off == ~0ull ? off : off + (unsigned long long)tot * 32ull;
		res->vbo_items = tot * 16u;
		res->ibo_items = tot * 6u;
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel 2: emission from the occupancy tile.
// ------------------------------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(kEmitThreads, RB == 7 ? 2 : 5)
k_mesh_emit(const VpWorldDev w, const uint32_t *__restrict__ ids, const uint8_t *__restrict__ scratch, size_t arrived_bytes, uint32_t cap_chunks,
            uint8_t *__restrict__ arena)
{
	using G = MGeo<RB>;
	constexpr int R = G::R, ZS = G::ZS, CL = G::CL;
	constexpr uint32_t FULL = 0xffffffffu;
	extern __shared__ __align__(128) uint8_t smem[];
	uint64_t *occ = reinterpret_cast<uint64_t *>(smem + G::E_OFF_OCC);
	uint32_t *gpre = reinterpret_cast<uint32_t *>(smem + G::E_OFF_GP);
	const uint16_t *gne = reinterpret_cast<const uint16_t *>(gpre + G::NG + 1);
	uint32_t *lut = reinterpret_cast<uint32_t *>(smem + G::E_OFF_LUT);
	uint4 *ent = reinterpret_cast<uint4 *>(smem + G::E_OFF_ENT);
	uint16_t *tbl = reinterpret_cast<uint16_t *>(smem + G::E_OFF_TBL);
	EMisc *misc = reinterpret_cast<EMisc *>(smem + G::E_OFF_MISC);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t slab = blockIdx.x;
	const uint32_t chunk_i = slab / CL;
	const int crank = CL > 1 ? (int)(slab % CL) : 0;
	const MeshScratch sc = carve_scratch<RB>(const_cast<uint8_t *>(scratch), arrived_bytes, cap_chunks);
	const uint32_t *rec = sc.rec + (size_t)slab * kMeshRec;
	const uint32_t faces = rec[0];
	if (faces == 0) return;
	const unsigned long long choff = (unsigned long long)rec[4] | ((unsigned long long)rec[5] << 32);
	if (choff == ~0ull) return;                               // arena overflow: nothing was reserved
	const uint32_t base = rec[1], total = rec[2];
	const uint32_t cid = ids[chunk_i];
	const int cx = (int)(cid & ((1u << w.bits[0]) - 1)), cy = (int)((cid >> w.bits[0]) & ((1u << w.bits[1]) - 1));
	const int cz = (int)(cid >> (w.bits[0] + w.bits[1]));
	if (tid == 0) {
		mbar_init(&misc->bar, 1);
		mbar_fence_init();
		mbar_arrive_expect_tx(&misc->bar, G::OCC_STRIDE * 8 + G::GP_STRIDE * 4);
		tma_load_1d(occ, sc.occ + (size_t)slab * G::OCC_STRIDE, G::OCC_STRIDE * 8, &misc->bar);
		tma_load_1d(gpre, sc.gp + (size_t)slab * G::GP_STRIDE, G::GP_STRIDE * 4, &misc->bar);
	}
	if (tid < 4) misc->slot[tid] = chunk_slot(w, cx + (tid == 1), cy + (tid == 2), cz + (tid == 3));
	fill_spread_lut(lut, tid, kEmitThreads);
	__syncthreads();
	if (tid < 32) mbar_wait(&misc->bar, 0);               // one warp polls the bulk copies, the others sleep in the barrier
	__syncthreads();

	const int z0 = crank * ZS;
	const MCtx<RB> mc{w, occ, misc->slot, z0, (uint32_t)cx << RB, (uint32_t)cy << RB, (uint32_t)cz << RB};
	uint4 *vbo = reinterpret_cast<uint4 *>(arena + choff);
	uint2 *ibo = reinterpret_cast<uint2 *>(arena + choff + (unsigned long long)total * 32ull);

	for (int ps = 0; ps < G::NPASS; ps++) {
		const int g0 = ps * G::PG;
		const uint32_t sbase = gpre[g0], S = gpre[g0 + G::PG] - sbase, ebase = gne[g0], n_ent = (uint32_t)gne[g0 + G::PG] - ebase;
		if (S == 0) continue;                                 // uniform: no face in these slices
		// ---- pass B: entries of the non-empty units, round table ----
		for (int g = warp; g < G::PG; g += kEmitWarps) {
			const uint32_t pg = gpre[g0 + g];
			if (gpre[g0 + g + 1] == pg) continue;
			const int u = (g0 + g) * 32 + lane;
			const uint64_t word = mc.unit_faces(lut, u);
			const uint32_t lo = (uint32_t)word, hi = (uint32_t)(word >> 32);
			const uint32_t c = __popc(lo) + __popc(hi);
			uint32_t inc = c;
			#pragma unroll
			for (int e = 1; e < 32; e <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, e); if (lane >= e) inc += t; }
			const uint32_t ne = __ballot_sync(FULL, c != 0u);
			if (c) {
				const uint32_t p = pg - sbase + inc - c;
				const uint32_t k = (uint32_t)gne[g0 + g] - ebase + (uint32_t)__popc(ne & ((1u << lane) - 1u));
				ent[k] = make_uint4(p, (uint32_t)u, lo, hi);
				for (uint32_t r = (p + 31u) >> 5; (r << 5) < p + c; r++) tbl[r] = (uint16_t)k;
			}
		}
		__syncthreads();
		// ---- rounds of 32 consecutive faces ----
		const uint32_t rounds = (S + 31u) >> 5;
		for (uint32_t r = (uint32_t)warp; r < rounds; r += kEmitWarps) {
			const uint32_t s0 = r << 5;
			const uint32_t i0 = tbl[r], j = i0 + (uint32_t)lane;
			const uint32_t pj = j < n_ent ? ent[j].x : 0xFFFFFFFFu;
			const uint32_t dd = pj - s0;
			const uint32_t heads = __reduce_or_sync(FULL, (lane != 0 && dd < 32u) ? (1u << dd) : 0u);
			const uint32_t t = min((uint32_t)lane, S - 1u - s0);
			const uint32_t rel = (uint32_t)__popc(heads & (0xFFFFFFFFu >> (31u - t)));
			const uint4 e = ent[i0 + rel];
			const int bit = select64m(e.z, e.w, __popc(e.z), s0 + t - e.x);
			if (s0 + (uint32_t)lane < S) {
				const int u = (int)e.y, jx = u % G::UPR, rr = u / G::UPR;
				const int x = jx * 16 + bit / 3, axis = bit % 3, y = rr & (R - 1), z = z0 + (rr >> RB);
				const uint32_t rank = base + sbase + s0 + (uint32_t)lane;
				mc.emit_face(vbo + (size_t)rank * 2, ibo + (size_t)rank * 3, rank, x, y, z, axis);
			}
		}
		__syncthreads();                                      // the next pass reuses ent / tbl
	}
}

template <int RB>
cudaError_t launch(const VpWorldDev &w, const uint32_t *d_ids, uint32_t n, VpResultDev *d_results, const uint32_t *d_result_pos,
                   uint8_t *arena, VpArenaDev *state, uint8_t *scratch, uint32_t scratch_chunks, cudaStream_t s)
{
	using G = MGeo<RB>;
	cudaError_t e = cudaFuncSetAttribute(k_mesh_count<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(k_mesh_emit<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::E_SMEM);
	if (e != cudaSuccess) return e;
	const size_t ab = arrived_region_bytes(scratch_chunks);
	k_mesh_count<RB><<<n * G::CL, kThreads, G::SMEM, s>>>(w, d_ids, scratch, ab, scratch_chunks, d_results, d_result_pos, state);
	k_mesh_emit<RB><<<n * G::CL, kEmitThreads, G::E_SMEM, s>>>(w, d_ids, scratch, ab, scratch_chunks, arena);
	return cudaGetLastError();
}

template <int RB> size_t scratch_bytes(uint32_t n)
{
	using G = MGeo<RB>;
	const size_t slabs = (size_t)n * G::CL;
	return arrived_region_bytes(n) + slabs * ((size_t)G::OCC_STRIDE * 8 + (size_t)G::GP_STRIDE * 4 + kMeshRec * 4) + 256;
}

} // namespace

cudaError_t vp_launch_mesh(const VpWorldDev &w, const uint32_t *d_ids, uint32_t n, VpResultDev *d_results, const uint32_t *d_result_pos,
                           uint8_t *arena, VpArenaDev *state, uint8_t *scratch, uint32_t scratch_chunks, cudaStream_t s)
{
	if (n == 0) return cudaSuccess;
	if (n > scratch_chunks) return cudaErrorInvalidValue;
	switch (w.rb) {
	case 4: return launch<4>(w, d_ids, n, d_results, d_result_pos, arena, state, scratch, scratch_chunks, s);
	case 5: return launch<5>(w, d_ids, n, d_results, d_result_pos, arena, state, scratch, scratch_chunks, s);
	case 6: return launch<6>(w, d_ids, n, d_results, d_result_pos, arena, state, scratch, scratch_chunks, s);
	case 7: return launch<7>(w, d_ids, n, d_results, d_result_pos, arena, state, scratch, scratch_chunks, s);
	default: return cudaErrorInvalidValue;
	}
}

// Bytes of device scratch for mesh rebuilds of up to n chunks per launch (arrival counters, which must be zero before
// the first launch, then occupancy tile + prefixes + record of every slab).
size_t vp_mesh_scratch_bytes(int rb, uint32_t n)
{
	switch (rb) { case 4: return scratch_bytes<4>(n); case 5: return scratch_bytes<5>(n); case 6: return scratch_bytes<6>(n); case 7: return scratch_bytes<7>(n); }
	return 0;
}
