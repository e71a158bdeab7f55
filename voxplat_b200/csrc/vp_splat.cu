// vp_splat.cu -- cull + 5-level LOD + splat-list emission for a batch of chunks: two kernels, no communication between
// CTAs.
//
// Replaces, byte for byte, the splat branch of the reference dispatcher (chunkset.c:371-458):
//   chunk_make_mask (mesher.c:377-456) -> chunk_make_splatlist level 0 (mesher.c:497-536)
//   -> 4 x { chunk_mask_downsample (mesher.c:460-493) ; chunk_make_splatlist }
//
// k_splat_count -- one CTA of 8 warps per 16-slice z-slab of a chunk:
//   1. every warp streams its share of the slab's 2 KB tiles (+ the slice below and above) through its own slot of a
//      shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier) and refills the slot itself -- no producer
//      warp, no polling; the +x / +y halo bytes of the neighbour chunks come by further bulk copies;
//   2. packs bytes into 64-bit occupancy rows (bit x of row (z,y)): SWAR non-zero flags gathered with IDP.4A;
//   3. derives the visibility rows with shift / AND-NOT face tests, then the 4 LOD levels by pair-OR-compress of the
//      bit rows -- no byte mask is ever materialised; the +x plane bits are also scattered into a bit string indexed by
//      row word ("unit"), so that counting and emission read them with one shift instead of a division;
//   4. counts groups of 32 row words with popc + REDUX, scans the group counts (stable z,y,x order comes from prefixes);
//   5. bulk-copies the bit arrays + group prefixes to a scratch in HBM; the LAST slab of a chunk to finish (arrival
//      counter) reserves the chunk's contiguous [L0|L1|L2|L3|L4] buffer with one atomicAdd and writes the slab bases.
// k_splat_emit -- one CTA of 8 warps per slab, 64 warps per SM: fetches the slab's bit arrays with bulk copies; warps take
//   non-empty groups from a ticket; slot -> unit by popc over the window's start mask, slot -> bit by popc halving + a
//   table; position from the bit index, colour byte gathered from the voxels (for LOD >= 1 the colour of the "last
//   non-zero child in scan order", found by descending the bit pyramids), shadow bit from two uint16 loads.
// (A fused single-kernel variant -- slab records staged in L2, last slab of a chunk places them -- was built and measured
// this round: byte-identical, 15 % less DRAM traffic, but 0.43-0.47 ms against this pair's time: its 90-150 KB of code
// stalled on instruction fetch and its barriers idled the warps.  DESIGN.md section 6.)
//
// The VP_* macros below are tuning hooks (scripts/build_variant.sh); the defaults are the measured best for 64^3 chunks.
#include "vp_device.cuh"
#include <cstddef>
using namespace vp;

namespace {

#ifndef VP_PREFETCH
#define VP_PREFETCH 0                 // L2 prefetch of a warp's later tiles: measured slower (the ring refills keep 16 KB per CTA in flight)
#endif
#ifndef VP_EMIT_STATIC
#define VP_EMIT_STATIC 0
#endif
#ifndef VP_EMIT_MINB
#define VP_EMIT_MINB 8
#endif
#ifndef VP_MINB6
#define VP_MINB6 7
#endif

template <int RB> struct Geo {
	static constexpr int R = 1 << RB;
	static constexpr int THREADS = 256;                  // count kernel
	static constexpr int NWARPS = THREADS / 32;          // every warp owns one slot of the TMA ring and refills it itself
	static constexpr int ZS = 16;                        // z slices per CTA
	static constexpr int CL = R / ZS;                    // CTAs (slabs) per chunk
	static constexpr int NW = R > 64 ? R / 64 : 1;       // 64-bit words per level-0 row
	static constexpr int SLICE = R * R;
	static constexpr int TILE = SLICE < 2048 ? SLICE : 2048;
	static constexpr int TPS = SLICE / TILE;             // tiles per slice
	static constexpr int RPT = TILE / R;                 // rows per tile
	static constexpr int NSL = ZS + 2;                   // slices streamed: z0-1 .. z0+ZS
	static constexpr int NT = NSL * TPS;
	static constexpr int LPR = R / 16;                   // lanes (16 B each) per row
	__host__ __device__ static constexpr int Rl(int l) { return R >> l; }
	__host__ __device__ static constexpr int Zl(int l) { return ZS >> l; }
	__host__ __device__ static constexpr int NWl(int l) { return Rl(l) > 64 ? Rl(l) / 64 : 1; }
	// level bit arrays (64-bit words): main[Zl][Rl+1][NWl] (row Rl = +y plane), xpl[Zl][NWl], zpl[Rl][NWl]
	__host__ __device__ static constexpr int main_words(int l) { return Zl(l) * (Rl(l) + 1) * NWl(l); }
	__host__ __device__ static constexpr int xpl_off(int l) { return main_words(l); }
	__host__ __device__ static constexpr int zpl_off(int l) { return main_words(l) + Zl(l) * NWl(l); }
	__host__ __device__ static constexpr int lvl_words(int l) { return zpl_off(l) + Rl(l) * NWl(l); }
	__host__ __device__ static constexpr int lvl_off(int l) { return l == 0 ? 0 : lvl_off(l - 1) + lvl_words(l - 1); }
	static constexpr int LV_WORDS = lvl_off(5);
	// rows in emission order: Zl*(Rl+1) slab rows, then Rl rows of the +z plane (top CTA only)
	__host__ __device__ static constexpr int nrows(int l) { return Zl(l) * (Rl(l) + 1) + Rl(l); }
	// "units" = the 64-bit words of the rows in emission order; groups of 32 units for counting / emission
	__host__ __device__ static constexpr int nunits(int l) { return nrows(l) * NWl(l); }
	__host__ __device__ static constexpr int ngroups(int l) { return (nunits(l) + 31) / 32; }
	__host__ __device__ static constexpr int grp_off(int l) { return l == 0 ? 0 : grp_off(l - 1) + ngroups(l - 1); }
	static constexpr int NG = grp_off(5);
	// count kernel shared memory (bytes): ring | halo | occ | occx, then barriers and the group record.  The level bit
	// arrays lv are built over the drained ring.
	static constexpr int RING_BYTES = NWARPS * TILE;
	static constexpr int LV_STRIDE = (LV_WORDS + 1) / 2 * 2;            // uint64 words, 16-byte multiple for bulk copies
	static constexpr int SCRATCH0 = ((RING_BYTES > LV_STRIDE * 8 ? RING_BYTES : LV_STRIDE * 8) + 127) / 128 * 128;
	static constexpr int HALO_BYTES = 2 * ZS * R;
	static constexpr int OCC_WORDS = NSL * (R + 1) * NW;
	static constexpr int OCCX_WORDS = ZS * NW;
	static constexpr int OFF_HALO = SCRATCH0;
	static constexpr int OFF_OCC = OFF_HALO + HALO_BYTES;
	static constexpr int OFF_OCCX = OFF_OCC + OCC_WORDS * 8;
	static constexpr int OFF_LV = 0;
	static constexpr int OFF_BARS = OFF_OCCX + OCCX_WORDS * 8;
	static constexpr int OFF_MISC = OFF_BARS + (NWARPS + 2) * 8;
	// group record: [NG + 1] exclusive splat prefix, [1] number of non-empty groups, [NG bytes] their indices,
	// [NG + 4] the +x plane bit of every unit (32 units per word)
	static constexpr int XB_OFF = NG + 2 + (NG + 3) / 4;                // uint32 words into the record
	static constexpr int GP_WORDS = XB_OFF + NG + 4;
	static constexpr int GP_STRIDE = (GP_WORDS + 3) / 4 * 4;            // uint32 words, 16-byte multiple for bulk copies
	static constexpr int SMEM = OFF_MISC + 48 + GP_STRIDE * 4 + 16;
	// emit kernel: level bit arrays | select table | group record | per-warp staging | barrier + scalars
	static constexpr int E_OFF_LUT = LV_STRIDE * 8;
	static constexpr int E_OFF_GP = E_OFF_LUT + 2048;
	static constexpr int E_OFF_STAGE = E_OFF_GP + GP_STRIDE * 4;        // per warp: 2 x 32 uint4 unit descriptors
	static constexpr int E_OFF_MISC = E_OFF_STAGE + 8 * 64 * 16;
	static constexpr int E_SMEM = E_OFF_MISC + 64;
	static_assert(NG + 4 <= THREADS, "xb bit string is zeroed by one pass of the CTA");
	static_assert(OFF_MISC % 16 == 0, "the group record is bulk-copied from 16-byte aligned shared memory");
};
constexpr int kEmitWarps = 8;      // E_OFF_MISC assumes 8 staging areas
constexpr int kChunkRec = 8;        // uint64 per chunk record
constexpr int kSlabRec = 16;        // uint32 per slab record: [0..4] counts, [8..12] first splat of the slab's part of level l

struct Misc {                       // count kernel: small per-CTA area in shared memory
	const uint8_t *src[4];          // own voxels / +x face plane / +y voxels / +z voxels of the chunk (null = null chunk)
	uint32_t pad[4];
	uint32_t gpre[1];               // the group record (extends past the struct): see Geo::GP_WORDS
};
static_assert(offsetof(Misc, gpre) == 48, "the group record starts 48 bytes into Misc");

enum CellKind { MAIN = 0, XPL = 1, YPL = 2, ZPL = 3 };

// kSelLut.v[b * 8 + k] = position of the k-th (0-based) set bit of byte b (8 where b has fewer bits).  Copied to
// shared memory by one bulk copy per CTA; finishes the rank->bit select of the emission with one LDS.
struct alignas(16) SelLut {
	uint8_t v[2048];
	constexpr SelLut() : v()
	{
		for (int b = 0; b < 256; b++)
			for (int k = 0; k < 8; k++) {
				int n = 0, pos = 8;
				for (int i = 0; i < 8; i++)
					if (b >> i & 1) { if (n == k) { pos = i; break; } n++; }
				v[b * 8 + k] = (uint8_t)pos;
			}
	}
};
__device__ const SelLut kSelLut = SelLut();

// Optional phase timing (VP_NVCC_EXTRA=-DVP_PROFILE_PHASES, scripts/phase_probe.py): thread 0 of every CTA
// accumulates clock64 deltas per phase.  Compiled out by default.
#ifdef VP_PROFILE_PHASES
__device__ unsigned long long g_phase_cycles[8];
#define VP_PHASE(k) do { if (threadIdx.x == 0) { long long t__ = clock64(); atomicAdd(&g_phase_cycles[k], (unsigned long long)(t__ - t_prev__)); t_prev__ = t__; } } while (0)
#define VP_PHASE_INIT long long t_prev__ = clock64()
#else
#define VP_PHASE(k) do { } while (0)
#define VP_PHASE_INIT do { } while (0)
#endif

template <int RB> struct Ctx {
	using G = Geo<RB>;
	const VpWorldDev &w;
	const uint64_t *lv;             // level bit arrays
	const uint8_t *own, *nbx_xlo, *nby, *nbz;     // chunk bytes (may be null)
	int z0;
	uint32_t ox, oy, oz;            // chunk origin in world voxels
	const uint8_t *own_z, *nby_z;   // own / +y neighbour voxels at the slab's first slice
	uint32_t xo_z;                  // offset of the slab's first row in the +x face plane

	__device__ __forceinline__ static uint32_t pair_at(const uint64_t *row, int bit) { return (uint32_t)(row[bit >> 6] >> (bit & 63)) & 3u; }

	// One step of the descent from level K to K-1: take the highest-priority set child, i.e. the LAST
	// non-zero child in (z,y,x) scan order (mesher.c:474-490).  All array offsets are compile-time.
	template <int K>
	__device__ __forceinline__ void descend(int kind, int &X, int &Y, int &Z) const
	{
		if constexpr (K >= 1) {
			constexpr int c = K - 1, Rc = G::Rl(c), NWc = G::NWl(c);
			const uint64_t *cm = lv + G::lvl_off(c);
			if (kind == MAIN) {
				// the four child rows' bit pairs, packed in scan order (dz, dy, dx): the highest set bit is the last
				// non-zero child -- four independent loads instead of a chain of tests
				const uint64_t *r0 = cm + ((2 * Z) * (Rc + 1) + 2 * Y) * NWc;
				const int wsel = (2 * X) >> 6, sft = (2 * X) & 63;
				const uint32_t p00 = (uint32_t)(r0[wsel] >> sft) & 3u, p01 = (uint32_t)(r0[NWc + wsel] >> sft) & 3u;
				const uint32_t p10 = (uint32_t)(r0[(Rc + 1) * NWc + wsel] >> sft) & 3u, p11 = (uint32_t)(r0[(Rc + 2) * NWc + wsel] >> sft) & 3u;
				const int top = 31 - __clz((int)(p00 | (p01 << 2) | (p10 << 4) | (p11 << 6)));
				Z = 2 * Z + (top >> 2); Y = 2 * Y + ((top >> 1) & 1); X = 2 * X + (top & 1);
			} else if (kind == XPL) {
				uint32_t p = pair_at(cm + G::xpl_off(c) + (2 * Z + 1) * NWc, 2 * Y);
				if (p) { Z = 2 * Z + 1; } else { p = pair_at(cm + G::xpl_off(c) + (2 * Z) * NWc, 2 * Y); Z = 2 * Z; }
				Y = 2 * Y + (int)(p >> 1);
			} else if (kind == YPL) {
				uint32_t p = pair_at(cm + ((2 * Z + 1) * (Rc + 1) + Rc) * NWc, 2 * X);
				if (p) { Z = 2 * Z + 1; } else { p = pair_at(cm + ((2 * Z) * (Rc + 1) + Rc) * NWc, 2 * X); Z = 2 * Z; }
				X = 2 * X + (int)(p >> 1);
			} else {
				uint32_t p = pair_at(cm + G::zpl_off(c) + (2 * Y + 1) * NWc, 2 * X);
				if (p) { Y = 2 * Y + 1; } else { p = pair_at(cm + G::zpl_off(c) + (2 * Y) * NWc, 2 * X); Y = 2 * Y; }
				X = 2 * X + (int)(p >> 1);
			}
			descend<K - 1>(kind, X, Y, Z);
		}
	}

	// Colour byte of the level-L cell x of row q (rows in emission order): the voxel found by descending the bit
	// pyramids to the last non-zero child, one byte gather (L2: this CTA streamed the voxel moments ago).
	template <int L>
	__device__ __forceinline__ uint32_t colour(int q, int x) const
	{
		constexpr int R = G::R, Rl = G::Rl(L), n_main = G::Zl(L) * (Rl + 1);
		int kind, cxx = x, cyy, czz;
		if (q < n_main) { czz = q / (Rl + 1); cyy = q - czz * (Rl + 1); kind = cyy == Rl ? YPL : (x == Rl ? XPL : MAIN); }
		else { kind = ZPL; cyy = q - n_main; czz = 0; }
		descend<L>(kind, cxx, cyy, czz);
		const uint8_t *p;
		if (kind == MAIN) p = own + ((size_t)(z0 + czz) * R + cyy) * R + cxx;
		else if (kind == XPL) p = nbx_xlo + (size_t)(z0 + czz) * R + cyy;
		else if (kind == YPL) p = nby + (size_t)(z0 + czz) * R * R + cxx;
		else p = nbz + (size_t)cyy * R + cxx;
		return __ldg(p);
	}
};

// OR the low `nbits` (<= 64) bits of `v` into a bit string of 32-bit words at bit offset `bit` (shared-memory atomics: the
// pieces of neighbouring writers share words).
__device__ __forceinline__ void scatter_or(uint32_t *bs, uint32_t bit, uint64_t v)
{
	const uint32_t w = bit >> 5, sh = bit & 31u;
	const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
	const uint32_t a = lo << sh;
	const uint32_t b = sh ? ((lo >> (32u - sh)) | (hi << sh)) : hi;
	const uint32_t c = sh ? (hi >> (32u - sh)) : 0u;
	if (a) atomicOr(bs + w, a);
	if (b) atomicOr(bs + w + 1, b);
	if (c) atomicOr(bs + w + 2, c);
}

// spread the 32 bits of v to the odd bit positions of a 64-bit value (bit i -> bit 2 i + 1)
__device__ __forceinline__ uint64_t spread_odd(uint32_t v)
{
	uint64_t x = v;
	x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
	x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
	x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
	x = (x | (x << 2)) & 0x3333333333333333ull;
	x = (x | (x << 1)) & 0x5555555555555555ull;
	return x << 1;
}

// The 64-bit word of unit u of level L (u < nunits(L)).
template <int RB, int L>
__device__ __forceinline__ uint64_t unit_word(const uint64_t *lv, uint32_t u)
{
	using G = Geo<RB>;
	constexpr uint32_t n_mw = G::main_words(L);                          // == n_main * NWl
	return lv[G::lvl_off(L) + u + (u >= n_mw ? (uint32_t)(G::zpl_off(L) - G::main_words(L)) : 0u)];
}

// Unit u = 32 g + lane of level L: its word (0 past the level's last unit) and the +x plane bit that follows it in scan
// order: bit `lane` of word grp_off(L) + g of the xb bit string the visibility / LOD phases scattered the plane bits into.
template <int RB, int L>
__device__ __forceinline__ uint64_t load_unit(const uint64_t *lv, const uint32_t *xbs, int g, int lane, uint32_t &xb)
{
	using G = Geo<RB>;
	const uint32_t u = (uint32_t)(g * 32 + lane);
	xb = (xbs[G::grp_off(L) + g] >> lane) & 1u;
	return u < (uint32_t)G::nunits(L) ? unit_word<RB, L>(lv, u) : 0ull;
}

// Position of the k-th (0-based) set bit of hi:lo (k < popc): three popc halving steps down to one byte, then the
// shared-memory table.  cl = popc(lo).
__device__ __forceinline__ uint32_t select64(uint32_t lo, uint32_t hi, uint32_t cl, uint32_t k, const uint8_t *lut)
{
	uint32_t v = lo, c, pos = 0;
	if (k >= cl) { k -= cl; v = hi; pos = 32; }
	c = __popc(v & 0xFFFFu); if (k >= c) { k -= c; v >>= 16; pos += 16; }
	c = __popc(v & 0xFFu);   if (k >= c) { k -= c; v >>= 8;  pos += 8; }
	return pos + lut[(v & 0xFFu) * 8u + (k & 7u)];
}

// Emission of one group of 32 units of level L by one warp.  Lane i owns unit i: its word, its exclusive splat
// prefix and a small descriptor of the row the unit lies in (record halves, shadow index, source byte row).
// Every round the 32 lanes take 32 consecutive output slots; slot s belongs to the unit i with
// p_i <= s < p_i + c_i (5-step shuffle binary search over the prefixes) and, inside it, to its (s - p_i)-th set
// bit (select64); the +x plane cell that follows a slab row in scan order is the unit's last slot.  Everything
// the record needs from the unit comes over shuffles, so a slot costs no divisions and no 64-bit index math.
template <int RB, int L>
__device__ __forceinline__ void emit_group(const Ctx<RB> &cx, const uint64_t *lv, const uint32_t *xbs, const uint8_t *lut, uint4 *stage, int gl, uint2 *out_g, int lane)
{
	using G = Geo<RB>;
	constexpr int R = G::R, Rl = G::Rl(L), NWl = G::NWl(L), n_main = G::Zl(L) * (Rl + 1);
	constexpr uint32_t FULL = 0xffffffffu;
	constexpr uint32_t XW = Rl < 64 ? Rl : 64;           // x of the +x plane cell relative to the unit's first bit
	constexpr uint32_t d = L ? (1u << L) : 0u;           // shadow sample offset of LOD splats (mesher.c:526-531)
	const int u = gl * 32 + lane;
	uint32_t xb;
	const uint64_t word = load_unit<RB, L>(lv, xbs, gl, lane, xb);
	const uint32_t lo = (uint32_t)word, hi = (uint32_t)(word >> 32);
	const uint32_t c = __popc(lo) + __popc(hi) + xb;
	uint32_t inc = c;
	#pragma unroll
	for (int e = 1; e < 32; e <<= 1) { uint32_t t = __shfl_up_sync(FULL, inc, e); if (lane >= e) inc += t; }
	const uint32_t p = inc - c, S = __shfl_sync(FULL, inc, 31);

	// unit descriptor
	const int q = u / NWl, xbase = (u % NWl) * 64;
	int Y, Zloc; uint32_t Zc;
	if (q < n_main) { Zloc = q / (Rl + 1); Y = q - Zloc * (Rl + 1); Zc = (uint32_t)((cx.z0 >> L) + Zloc); }
	else { Zloc = 0; Y = q - n_main; Zc = Rl; }
	const uint32_t wx0 = cx.ox + ((uint32_t)xbase << L), wy = cx.oy + ((uint32_t)Y << L), wz = cx.oz + (Zc << L);
	const uint32_t A = (wx0 & 0xFFFFu) | (wy << 16), B = wz & 0xFFFFu;       // world size <= 32768 per axis: wy fits 16 bits
	const uint32_t shb = (wx0 + d) + (wy + d) + cx.w.sh_w * (wz + d - cx.w.sh_z0);
	const uint8_t *row = nullptr; uint32_t xo = 0;
	if constexpr (L == 0) {
		if (q < n_main) {
			// q = Zloc * (R + 1) + Y, so the row's voxel offset inside the slab is (q - Zloc) * R; the slab base pointers
			// were computed once per CTA
			const bool yp = Y >= R;
			row = (yp ? cx.nby_z : cx.own_z) + ((yp ? (uint32_t)Zloc << (2 * RB) : (uint32_t)(q - Zloc) << RB) + (uint32_t)xbase);
			xo = cx.xo_z + (uint32_t)(q - Zloc);
		} else {
			row = cx.nbz + (size_t)Y * R + xbase;
		}
	}
	const unsigned long long rowa = (unsigned long long)row;
	const uint32_t rlo = (uint32_t)rowa, rhi = (uint32_t)(rowa >> 32);
	const uint32_t Bx = B | (xo << 16);                  // xo < R*R + R <= 16512: both halves fit 16 bits

	// The non-empty units are compacted into the warp's staging rows (rank = popc of the ballot below the lane): their
	// prefixes are then strictly increasing, so "which unit owns slot s" is a popc over the mask of unit starts inside
	// the round's window instead of a shuffle binary search, and the 8 descriptor words come back as two 16-byte reads.
	const uint32_t ne = __ballot_sync(FULL, c != 0u), n_ne = __popc(ne);
	const uint32_t rank = __popc(ne & ((1u << lane) - 1u));
	__syncwarp();                                        // the previous group's readers are done with the staging rows
	if (c) {
		stage[rank] = make_uint4(p, lo, hi, A);
		stage[32 + rank] = make_uint4(Bx, shb, L == 0 ? rlo : (uint32_t)u, rhi);
	}
	__syncwarp();
	const uint32_t pc = (uint32_t)lane < n_ne ? stage[lane].x : 0xFFFFFFFFu;       // start slot of compacted unit `lane`

	for (uint32_t s0 = 0; s0 < S; s0 += 32) {
		const uint32_t t = min((uint32_t)lane, S - 1 - s0), s = s0 + t;
		const uint32_t dd = pc - s0;
		const uint32_t heads = __reduce_or_sync(FULL, dd < 32u ? 1u << dd : 0u);       // unit starts inside [s0, s0 + 32)
		const uint32_t before = __popc(__ballot_sync(FULL, pc < s0));                  // units that start before the window
		const uint32_t i = before - 1u + __popc(heads & (0xFFFFFFFFu >> (31u - t)));
		const uint4 d0 = stage[i], d1 = stage[32 + i];
		const uint32_t pi = d0.x, wlo = d0.y, whi = d0.z, uA = d0.w, uB = d1.x, ush = d1.y;
		const uint32_t k = s - pi, wcl = __popc(wlo), wcw = wcl + __popc(whi);
		const bool isx = k >= wcw;                       // the unit's +x plane cell (only ever its last slot)
		const uint32_t pos = isx ? XW : select64(wlo, whi, wcl, k, lut);
		const uint32_t xs = pos << L;
		// shadow_sample (shadow.h:56-63): !(map[idx] < y+1 && map[idx+1] < y+1)
		const uint32_t lim = (uA >> 16) + d + 1u;
		const uint16_t *sp = cx.w.shadow + (ush + xs);
		const uint32_t sa = __ldg(sp), sb = __ldg(sp + 1);        // both loads in flight together
		const uint32_t sh = ((sa >= lim) | (sb >= lim)) ? 64u : 0u;
		uint32_t col;
		if constexpr (L == 0) {
			const uint8_t *src = isx ? cx.nbx_xlo + (uB >> 16) : reinterpret_cast<const uint8_t *>(((unsigned long long)d1.w << 32) | d1.z) + pos;
			col = __ldg(src);
		} else {
			const int uu = (int)d1.z;                    // unit index inside the level
			col = cx.template colour<L>(uu / NWl, (int)pos + 64 * (uu % NWl));
		}
		if (s0 + lane < S) {
			const uint32_t rl = ((uA + xs) & 0xFFFFu) | (uA & 0xFFFF0000u);
			const uint32_t rh = __byte_perm(uB, col | sh, 0x5410);       // wz | (colour | shadow) << 16
			out_g[s0 + lane] = make_uint2(rl, rh);
		}
	}
}

// Scratch between the two kernels of a splat rebuild (device pointers, sized by vp_splat_scratch_bytes).
struct SplatScratch {
	uint32_t *arrived;              // [cap chunks]        slabs of the chunk that finished counting (self-resetting)
	uint64_t *pyr;                  // [slabs][LV_STRIDE]  level bit arrays of every non-empty slab
	uint32_t *gp;                   // [slabs][GP_STRIDE]  exclusive prefix of the per-group splat counts
	uint32_t *rec;                  // [slabs][kSlabRec]   per-level counts (count kernel) and bases (scan kernel)
	unsigned long long *chrec;      // [chunks][kChunkRec] [0] byte offset of the chunk buffer in the arena (~0 = none),
	                                //                     [1..4] own / +x face plane / +y / +z voxel pointers (0 = null chunk)
};

template <int RB>
__host__ __device__ __forceinline__ SplatScratch carve_scratch(uint8_t *base, size_t arrived_bytes, uint32_t n)
{
	using G = Geo<RB>;
	const size_t slabs = (size_t)n * G::CL;
	SplatScratch sc;
	sc.arrived = reinterpret_cast<uint32_t *>(base);                    // fixed position whatever n: stays zero between launches
	base += arrived_bytes;
	sc.pyr = reinterpret_cast<uint64_t *>(base);
	sc.gp = reinterpret_cast<uint32_t *>(base + slabs * G::LV_STRIDE * 8);
	sc.rec = sc.gp + slabs * G::GP_STRIDE;
	sc.chrec = reinterpret_cast<unsigned long long *>(sc.rec + slabs * kSlabRec);
	return sc;
}

// End of a slab's count pass: the LAST slab of a chunk to get here (per-chunk arrival counter) adds up the level
// counts of all slabs, reserves the chunk's contiguous [L0|L1|L2|L3|L4] buffer in the arena with one atomicAdd, and
// writes the per-slab level bases and the result record (ChunkMD.svl_items[], chunkset.c:469-483).  Nobody waits.
template <int CL>
__device__ __forceinline__ void chunk_reserve(const SplatScratch &sc, uint32_t chunk_i, VpResultDev *res, VpArenaDev *st)
{
	__syncthreads();                                   // this slab's record is written
	if (threadIdx.x >= 32) return;
	const int lane = threadIdx.x;
	if (CL > 1) {
		uint32_t prev = 0;
		if (lane == 0) { __threadfence(); prev = atomicAdd(sc.arrived + chunk_i, 1u); }
		prev = __shfl_sync(0xffffffffu, prev, 0);
		if (prev != CL - 1) return;
		if (lane == 0) sc.arrived[chunk_i] = 0;        // ready for the next launch
		__threadfence();
	}
	// lane = 8 * (slab within the pass) + level: all records are read at once, sums and prefixes by shuffles
	uint32_t *rc = sc.rec + (size_t)chunk_i * CL * kSlabRec;
	const int l = lane & 7, rr = lane >> 3;
	constexpr int PASSES = (CL + 3) / 4;
	uint32_t cnt[PASSES], incl[PASSES], tot = 0;
	#pragma unroll
	for (int p = 0; p < PASSES; p++) {
		const int r = p * 4 + rr;
		cnt[p] = (l < 5 && r < CL) ? __ldcg(rc + r * kSlabRec + l) : 0u;
		uint32_t x = cnt[p], t;
		t = __shfl_up_sync(0xffffffffu, x, 8);  if (rr >= 1) x += t;
		t = __shfl_up_sync(0xffffffffu, x, 16); if (rr >= 2) x += t;
		incl[p] = tot + x;                             // slabs 0..r of level l
		tot += __shfl_sync(0xffffffffu, x, 24 + l);    // + this pass's 4 slabs
	}
	// exclusive prefix of the level totals over l (lanes of one 8-lane group)
	uint32_t pre = tot, t;
	t = __shfl_up_sync(0xffffffffu, pre, 1, 8); if (l >= 1) pre += t;
	t = __shfl_up_sync(0xffffffffu, pre, 2, 8); if (l >= 2) pre += t;
	t = __shfl_up_sync(0xffffffffu, pre, 4, 8); if (l >= 4) pre += t;
	const uint32_t base_l = pre - tot;
	const uint32_t total = __shfl_sync(0xffffffffu, pre, 4);               // levels 0..4
	#pragma unroll
	for (int p = 0; p < PASSES; p++) {
		const int r = p * 4 + rr;
		if (l < 5 && r < CL) rc[r * kSlabRec + 8 + l] = base_l + incl[p] - cnt[p];
	}
	unsigned long long off = 0;
	if (lane == 0 && total) {
		const unsigned long long bytes = (unsigned long long)total * 8ull;
		off = atomicAdd(&st->cursor, bytes);
		if (off + bytes > st->capacity) { atomicExch(&st->overflow, 1u); off = ~0ull; }
	}
	if (lane == 0) {
		sc.chrec[(size_t)chunk_i * kChunkRec] = total ? off : ~0ull;
		res->svl_offset = off;
		res->svl_items_total = total * 4u;
	}
	if (lane < 5) res->svl_items[lane] = tot * 4u;
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel 1: stream one 16-slice slab of a chunk, pack it to bits, derive visibility + LOD bit arrays, count.
// One CTA per slab, no communication between CTAs: the bit arrays, the group prefixes and the 5 level counts go to
// the scratch, the scan kernel turns the counts of all slabs into arena offsets, the emit kernel writes the splats.
// ------------------------------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(Geo<RB>::THREADS, RB == 6 ? VP_MINB6 : (RB < 6 ? 5 : 2))
k_splat_count(const VpWorldDev w, const uint32_t *__restrict__ ids, uint32_t n, uint8_t *__restrict__ scratch, size_t arrived_bytes,
              VpResultDev *__restrict__ results, const uint32_t *__restrict__ result_pos, VpArenaDev *__restrict__ st)
{
	using G = Geo<RB>;
	constexpr int R = G::R, ZS = G::ZS, CL = G::CL, NW = G::NW, TILE = G::TILE, TPS = G::TPS, NT = G::NT;
	constexpr int kThreads = G::THREADS, kWarps = G::NWARPS;
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t *ring = smem;
	uint8_t *halo = smem + G::OFF_HALO;
	uint64_t *occ = reinterpret_cast<uint64_t *>(smem + G::OFF_OCC);
	uint64_t *occx = reinterpret_cast<uint64_t *>(smem + G::OFF_OCCX);
	uint64_t *lv = reinterpret_cast<uint64_t *>(smem + G::OFF_LV);
	uint64_t *bar_slot = reinterpret_cast<uint64_t *>(smem + G::OFF_BARS);       // [NWARPS] one per ring slot = per warp
	uint64_t *bar_halo = bar_slot + kWarps;
	Misc *misc = reinterpret_cast<Misc *>(smem + G::OFF_MISC);
	uint32_t *gpre = misc->gpre;
	uint32_t *xbs = gpre + G::XB_OFF;                                            // [NG + 4] +x plane bit of every unit, 32 units per word

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int crank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
	const uint32_t chunk_i = blockIdx.x / CL;
	const int z0 = crank * ZS;
	const bool top = (z0 + ZS == R);
	const SplatScratch sc = carve_scratch<RB>(scratch, arrived_bytes, n);
	uint32_t *rec = sc.rec + (size_t)blockIdx.x * kSlabRec;
	VpResultDev *res = results + (result_pos ? result_pos[chunk_i] : chunk_i);

	VP_PHASE_INIT;
	// ---- phase 0: source pointers (4 threads, one slot lookup each), barriers, zeroed bit arrays -------
	if (tid < 4) {
		const uint32_t cid = ids[chunk_i];
		const int cx = (int)(cid & ((1u << w.bits[0]) - 1)), cy = (int)((cid >> w.bits[0]) & ((1u << w.bits[1]) - 1));
		const int cz = (int)(cid >> (w.bits[0] + w.bits[1]));
		const int sl = chunk_slot(w, cx + (tid == 1), cy + (tid == 2), cz + (tid == 3));
		const uint8_t *ptr = nullptr;
		if (sl >= 0) ptr = tid == 1 ? w.xlo_pool + (size_t)sl * R * R : w.vox_pool + (size_t)sl * R * R * R;
		misc->src[tid] = ptr;
		if (crank == 0) sc.chrec[(size_t)chunk_i * kChunkRec + 1 + tid] = (unsigned long long)ptr;       // for the emit kernel
	}
	if (tid == 32) {
		for (int i = 0; i < kWarps; i++) mbar_init(bar_slot + i, 1);
		mbar_init(bar_halo, 1);
		mbar_fence_init();
	}
	for (int i = tid; i < G::OCC_WORDS + G::OCCX_WORDS; i += kThreads) occ[i] = 0;      // occ, occx are contiguous
	if (tid < G::NG + 4) xbs[tid] = 0;
	__syncthreads();
	const uint8_t *own = misc->src[0], *nbx_xlo = misc->src[1], *nby = misc->src[2], *nbz = misc->src[3];
	if (!own && !nbx_xlo && !nby && !nbz) {           // mesher.c:404-409: nothing can be visible
		if (tid < 5) rec[tid] = 0;
		chunk_reserve<CL>(sc, chunk_i, res, st);
		return;
	}

	VP_PHASE(0);
	// Source of this warp's k-th tile, t = warp + kWarps k (slice z = z0 - 1 + t / TPS, part t % TPS of the slice); nullptr =
	// air or not needed.  Inside the chunk the tiles of a warp are kWarps tiles apart (TILE divides the slice).
	static_assert(kWarps % TPS == 0 || TPS % kWarps == 0, "tiles of a warp: fixed part of the slice or whole slices apart");
	const uint8_t *own_w = own ? own + ((ptrdiff_t)(z0 - 1) * R * R + (ptrdiff_t)warp * TILE) : nullptr;
	auto tile_src = [&](int t) -> const uint8_t * {
		const int z = z0 - 1 + t / TPS;
		if (z < 0) return nullptr;                     // no -z test at z = 0 (pair walk starts at A, mesher.c:421)
		if (z < R) return own_w ? own_w + (size_t)(t - warp) * TILE : nullptr;
		return nbz ? nbz + (size_t)(t % TPS) * TILE : nullptr;                       // z == R: slice 0 of the +z neighbour
	};
	const bool have_halo = nbx_xlo || nby;
	auto load_tile = [&](void *dst, const void *src, uint64_t *bar) { tma_load_1d(dst, src, TILE, bar); };

	// ---- phase 1: every warp streams its share of the tiles through its own ring slot (1-D TMA bulk copies) and packs
	// the bytes to occupancy bits; lane 0 refills the slot as soon as the warp has drained it ------------------------
	uint32_t any_solid = 0;
	{
		uint8_t *slot = ring + warp * TILE;
		uint64_t *bar = bar_slot + warp;
		if (lane == 0) {
			if (warp == 0) {
				if (have_halo) {
					mbar_arrive_expect_tx(bar_halo, (nbx_xlo ? ZS * R : 0) + (nby ? ZS * R : 0));
					if (nbx_xlo) tma_load_1d(halo, nbx_xlo + (size_t)z0 * R, ZS * R, bar_halo);
					if (nby) for (int z = 0; z < ZS; z++) tma_load_1d(halo + ZS * R + z * R, nby + (size_t)(z0 + z) * R * R, R, bar_halo);
				}
			}
			const uint8_t *src = tile_src(warp);
			if (src) { mbar_arrive_expect_tx(bar, TILE); load_tile(slot, src, bar); }
#if VP_PREFETCH
			for (int t = warp + kWarps; t < NT; t += kWarps) {         // this warp's later tiles: into L2 now
				const uint8_t *ps = tile_src(t);
				if (ps) l2_prefetch(ps, TILE);
			}
#endif
		}
		uint32_t ph = 0;
		const uint8_t *cur = tile_src(warp);
		for (int t = warp; t < NT; t += kWarps) {
			const int s = t / TPS, part = t % TPS;
			const uint8_t *nxt = t + kWarps < NT ? tile_src(t + kWarps) : nullptr;
			if (cur) {
				mbar_wait(bar, ph & 1); ph++;
				const uint8_t *tb = slot;
				uint64_t *orow = occ + (size_t)(s * (R + 1) + part * G::RPT) * NW;
				if constexpr (TILE >= 1024 && R >= 64) {
					// 32 bytes per lane and iteration: two conflict-free 16-byte reads 512 bytes apart.  Lane pairs
					// exchange their 16-bit masks with one shuffle; the even lane stores the 32-bit word of the first
					// read, the odd lane the word of the second, so every lane stores once.
					const uint32_t psel = (lane & 1) ? 0x3276u : 0x5410u;
					uint32_t *o32 = reinterpret_cast<uint32_t *>(orow);
					#pragma unroll
					for (int off = lane * 16; off < TILE; off += 1024) {
						const uint4 qa = *reinterpret_cast<const uint4 *>(tb + off);
						const uint4 qb = *reinterpret_cast<const uint4 *>(tb + off + 512);
						// all-air shortcut: the occupancy rows are pre-zeroed
						if (!__any_sync(0xffffffffu, (qa.x | qa.y | qa.z | qa.w | qb.x | qb.y | qb.z | qb.w) != 0u)) continue;
						any_solid = 1u;
						const uint32_t v = __byte_perm(nz16x128(qa) >> 7, nz16x128(qb) >> 7, 0x5410);
						const uint32_t pv = __shfl_xor_sync(0xffffffffu, v, 1);
						const int o = (lane & 1) ? off + 512 : off;
						o32[o >> 5] = __byte_perm(v, pv, psel);          // rows of a slice are contiguous 32-bit words
					}
				} else {
					#pragma unroll 4
					for (int off = lane * 16; off < TILE; off += 512) {
						const uint4 q4 = *reinterpret_cast<const uint4 *>(tb + off);
						if (TILE >= 512 && !__any_sync(0xffffffffu, (q4.x | q4.y | q4.z | q4.w) != 0u)) continue;
						const uint32_t m = nz16(q4);
						any_solid |= m;
						const int row = off / R, bo = off % R;
						if (R >= 32) {
							uint32_t v = m << (bo & 16);
							v |= __shfl_xor_sync(0xffffffffu, v, 1);
							if (!(lane & 1)) reinterpret_cast<uint32_t *>(orow + row * NW)[bo >> 5] = v;
						} else {
							reinterpret_cast<uint32_t *>(orow + row * NW)[0] = m;
						}
					}
				}
			}
			__syncwarp();                                  // every lane is done with the slot
			if (lane == 0 && nxt) { mbar_arrive_expect_tx(bar, TILE); load_tile(slot, nxt, bar); }
			cur = nxt;
		}
		// halo rows: 0..ZS-1 = x = 0 column of the +x neighbour (bits over y), ZS..2ZS-1 = y = 0 row of +y
		if (have_halo) {
			mbar_wait(bar_halo, 0);
			for (int f0 = warp * 32; f0 < 2 * ZS * G::LPR; f0 += kThreads) {       // whole warps: the pair shuffle needs every lane
				const int f = f0 + lane;
				const bool valid = f < 2 * ZS * G::LPR;
				const int hr = valid ? f / G::LPR : 0, bo = (f % G::LPR) * 16;
				const bool isx = hr < ZS;
				const bool present = valid && (isx ? (nbx_xlo != nullptr) : (nby != nullptr));
				uint32_t m = present ? nz16(*reinterpret_cast<const uint4 *>(halo + hr * R + bo)) : 0u;
				any_solid |= m;
				uint64_t *dst = isx ? occx + hr * NW : occ + (size_t)((hr - ZS + 1) * (R + 1) + R) * NW;
				if (R >= 32) {
					uint32_t v = m << (bo & 16);
					v |= __shfl_xor_sync(0xffffffffu, v, 1);
					if (present && !(lane & 1)) reinterpret_cast<uint32_t *>(dst)[bo >> 5] = v;
				} else if (present) {
					reinterpret_cast<uint32_t *>(dst)[0] = m;
				}
			}
		}
	}
	// A slab without any solid voxel (own, slice above, +x/+y halo) has nothing visible.
	const bool nonempty = __syncthreads_or(any_solid != 0u) != 0;
	if (!nonempty) {
		if (tid < 5) rec[tid] = 0;
		chunk_reserve<CL>(sc, chunk_i, res, st);
		return;
	}
	uint64_t *lv0 = lv;                                // built over the drained ring; every word of it is written below

	// ---- phase 2: visibility rows (closed form of the pair walk, mesher.c:421-448) -------------------
	for (int f0 = warp * 32; f0 < ZS * R; f0 += kWarps * 32) {
		const int f = f0 + lane, zi = f >> RB, y = f & (R - 1), s = zi + 1;
		const uint64_t *o = occ + (size_t)(s * (R + 1) + y) * NW;
		const uint64_t xbit = (occx[zi * NW + (y >> 6)] >> (y & 63)) & 1ull;
		uint64_t *xp = lv0 + G::xpl_off(0);
		// 32 rows of air with no solid +x halo cell: nothing of them is visible
		{
			uint64_t any = xbit;
			#pragma unroll
			for (int k = 0; k < NW; k++) any |= o[k];
			if (!__any_sync(0xffffffffu, any != 0ull)) {
				#pragma unroll
				for (int k = 0; k < NW; k++) lv0[(zi * (R + 1) + y) * NW + k] = 0ull;
				if (R >= 64) { if (lane == 0) reinterpret_cast<uint32_t *>(xp + zi * NW)[y >> 5] = 0u; }
				else if (R == 32) { if (lane == 0) xp[zi] = 0ull; }
				else if ((lane & 15) == 0) xp[zi] = 0ull;
				continue;
			}
		}
		#pragma unroll
		for (int k = 0; k < NW; k++) {
			const uint64_t ow = o[k];
			const uint64_t nxt = (k + 1 < NW) ? (o[k + 1] & 1ull) : xbit;
			const uint64_t px = (ow >> 1) | (nxt << ((R - 1) & 63));
			const uint64_t mx = (ow << 1) | (k > 0 ? (o[k - 1] >> 63) : 1ull);
			const uint64_t py = o[NW + k];
			const uint64_t my = y > 0 ? o[k - NW] : ~0ull;
			const uint64_t pz = occ[(size_t)((s + 1) * (R + 1) + y) * NW + k];
			const uint64_t mz = (z0 + zi) > 0 ? occ[(size_t)((s - 1) * (R + 1) + y) * NW + k] : ~0ull;
			lv0[(zi * (R + 1) + y) * NW + k] = ow & ~(px & mx & py & my & pz & mz);
		}
		const uint32_t vx = (uint32_t)(xbit & ~(o[NW - 1] >> ((R - 1) & 63)));
		const uint32_t bal = __ballot_sync(0xffffffffu, vx & 1u);
		if (R >= 64) { if (lane == 0) reinterpret_cast<uint32_t *>(xp + zi * NW)[y >> 5] = bal; }
		else if (R == 32) { if (lane == 0) xp[zi] = (uint64_t)bal; }
		else if ((lane & 15) == 0) xp[zi] = (bal >> lane) & 0xFFFFu;
		// the same bits by unit number (the row's last word is followed by its +x plane cell in scan order)
		if (bal) {
			if (R >= 32) {
				if (lane == 0) {
					const uint32_t row0 = (uint32_t)(zi * (R + 1) + y);            // y = first of the 32 rows (lane 0)
					if (NW == 1) scatter_or(xbs, row0, (uint64_t)bal);
					else scatter_or(xbs, row0 * NW, spread_odd(bal));               // NW == 2: the odd units
				}
			} else if ((lane & 15) == 0) {
				scatter_or(xbs, (uint32_t)(zi * (R + 1)), (uint64_t)((bal >> lane) & 0xFFFFu));
			}
		}
	}
	const bool zplane = top && nbz;
	for (int f = tid; f < (ZS + R) * NW; f += kThreads) {
		const int r = f / NW, k = f % NW;
		if (r < ZS) {            // +y plane row of slice r: solid in the neighbour, air below it in this chunk
			const uint64_t *o = occ + (size_t)((r + 1) * (R + 1)) * NW;
			lv0[(r * (R + 1) + R) * NW + k] = o[R * NW + k] & ~o[(R - 1) * NW + k];
		} else {                 // +z plane row y = r - ZS (only the chunk's top slab has one)
			const int y = r - ZS;
			lv0[G::zpl_off(0) + y * NW + k] = zplane ? (occ[(size_t)((ZS + 1) * (R + 1) + y) * NW + k] & ~occ[(size_t)(ZS * (R + 1) + y) * NW + k]) : 0ull;
		}
	}
	__syncthreads();

	// ---- phase 3: LOD pyramids, level l from l-1 by OR of the child rows + pair-OR-compress.  Level 1 by the whole
	// CTA; levels 2..4 are a few dozen words: warp 0 builds them (and counts them) while the other warps count the
	// groups of levels 0 and 1 (phase 4) -------------------------------------------------------------------------
	auto build_level = [&](int l, int first, int stride) {
		const int c = l - 1;
		const int Rl = G::Rl(l), Zl = G::Zl(l), Rc = G::Rl(c), NWc = G::NWl(c);
		const uint64_t *cm = lv + G::lvl_off(c);
		uint64_t *pm = lv + G::lvl_off(l);
		const int n_main = Zl * (Rl + 1), n_all = n_main + Zl + Rl;
		for (int g = first; g < n_all; g += stride) {
			uint64_t a0 = 0, a1 = 0;
			uint64_t *dst;
			if (g < n_main) {
				const int Z = g / (Rl + 1), Y = g % (Rl + 1);
				for (int dz = 0; dz < 2; dz++) {
					if (Y < Rl) {
						for (int dy = 0; dy < 2; dy++) {
							const uint64_t *r = cm + (size_t)((2 * Z + dz) * (Rc + 1) + 2 * Y + dy) * NWc;
							a0 |= r[0]; if (NWc > 1) a1 |= r[NWc - 1];
						}
					} else {
						const uint64_t *r = cm + (size_t)((2 * Z + dz) * (Rc + 1) + Rc) * NWc;
						a0 |= r[0]; if (NWc > 1) a1 |= r[NWc - 1];
					}
				}
				dst = pm + g;
			} else if (g < n_main + Zl) {
				const int Z = g - n_main;
				const uint64_t *r = cm + G::xpl_off(c) + (size_t)(2 * Z) * NWc;
				a0 = r[0] | r[NWc]; if (NWc > 1) a1 = r[NWc - 1] | r[2 * NWc - 1];
				dst = pm + G::xpl_off(l) + Z;
			} else {
				const int Y = g - n_main - Zl;
				const uint64_t *r = cm + G::zpl_off(c) + (size_t)(2 * Y) * NWc;
				a0 = r[0] | r[NWc]; if (NWc > 1) a1 = r[NWc - 1] | r[2 * NWc - 1];
				dst = pm + G::zpl_off(l) + Y;
			}
			uint64_t p = pair_or_compress(a0);
			if (NWc > 1) p |= pair_or_compress(a1) << 32;
			*dst = p;                                      // NWl(l) == 1 for every l >= 1 (R <= 128)
			if (g >= n_main && g < n_main + Zl && p) scatter_or(xbs + G::grp_off(l), (uint32_t)((g - n_main) * (Rl + 1)), p);      // +x plane bits by unit number
		}
	};
	build_level(1, tid, kThreads);
	__syncthreads();

	// ---- phase 4: counts.  The bit rows of all levels are cut into groups of 32 "units" (one 64-bit word
	// plus, for the last word of a slab row, the +x plane bit that follows it in scan order).  One warp
	// per group: popc + REDUX gives the group's splat count, a ballot its number of non-empty units; one
	// warp then scans the packed pairs.  Stable (z,y,x) order follows from the prefix, not from atomics. ---
	auto count_group = [&](int g) {
		uint32_t xb, c;
		if (g < G::grp_off(1)) { c = __popcll(load_unit<RB, 0>(lv, xbs, g, lane, xb)) + xb; }
		else if (g < G::grp_off(2)) { c = __popcll(load_unit<RB, 1>(lv, xbs, g - G::grp_off(1), lane, xb)) + xb; }
		else if (g < G::grp_off(3)) { c = __popcll(load_unit<RB, 2>(lv, xbs, g - G::grp_off(2), lane, xb)) + xb; }
		else if (g < G::grp_off(4)) { c = __popcll(load_unit<RB, 3>(lv, xbs, g - G::grp_off(3), lane, xb)) + xb; }
		else { c = __popcll(load_unit<RB, 4>(lv, xbs, g - G::grp_off(4), lane, xb)) + xb; }
		c = __reduce_add_sync(0xffffffffu, c);
		if (lane == 0) gpre[g] = c;
	};
	if (warp == 0) {
		build_level(2, lane, 32); __syncwarp();
		build_level(3, lane, 32); __syncwarp();
		build_level(4, lane, 32); __syncwarp();
		for (int g = G::grp_off(2); g < G::NG; g++) count_group(g);
	} else {
		for (int g = warp - 1; g < G::grp_off(2); g += kWarps - 1) count_group(g);
	}
	__syncthreads();
	VP_PHASE(3);
	if (warp == 0) {
		constexpr int IPT = (G::NG + 31) / 32;
		uint32_t v[IPT], sum = 0;
		#pragma unroll
		for (int k = 0; k < IPT; k++) { const int g = lane * IPT + k; v[k] = g < G::NG ? gpre[g] : 0u; sum += v[k]; }
		uint32_t inc = sum;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
		uint32_t pre = inc - sum;
		#pragma unroll
		for (int k = 0; k < IPT; k++) { const int g = lane * IPT + k; if (g < G::NG) gpre[g] = pre; pre += v[k]; }
		if (lane == 31) gpre[G::NG] = pre;
		__syncwarp();
		constexpr int GO[6] = {G::grp_off(0), G::grp_off(1), G::grp_off(2), G::grp_off(3), G::grp_off(4), G::grp_off(5)};
		#pragma unroll
		for (int l = 0; l < 5; l++) if (lane == l) rec[l] = gpre[GO[l + 1]] - gpre[GO[l]];
		// list of the non-empty groups: the emit kernel's warps take entries of this list, not all NG groups
		static_assert(G::NG <= 255, "group indices are stored as bytes");
		uint8_t *glist = reinterpret_cast<uint8_t *>(gpre + G::NG + 2);
		uint32_t nne = 0;
		for (int g0 = 0; g0 < G::NG; g0 += 32) {
			const int g = g0 + lane;
			const bool ne = g < G::NG && gpre[g + 1] != gpre[g];
			const uint32_t m = __ballot_sync(0xffffffffu, ne);
			if (ne) glist[nne + __popc(m & ((1u << lane) - 1u))] = (uint8_t)g;
			nne += __popc(m);
		}
		if (lane == 0) gpre[G::NG + 1] = nne;
	}
	fence_proxy_async_smem();          // every thread: its writes to the bit arrays become visible to the bulk-copy engine
	__syncthreads();
	VP_PHASE(4);
	// ---- phase 5: bit arrays + group prefixes to the scratch (skipped when nothing is visible) -------
	if (gpre[G::NG] != 0 && tid == 0) {
		// two bulk copies shared -> global; the copy engine reads shared memory while the CTA goes on to the
		// reservation, thread 0 waits for those reads before it exits
		bulk_store(sc.pyr + (size_t)blockIdx.x * G::LV_STRIDE, lv, G::LV_STRIDE * 8);
		bulk_store(sc.gp + (size_t)blockIdx.x * G::GP_STRIDE, gpre, G::GP_STRIDE * 4);
		bulk_commit();
	}
	chunk_reserve<CL>(sc, chunk_i, res, st);
	if (tid == 0) bulk_wait_read();
	VP_PHASE(5);
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel 2: emission.  One CTA per slab fetches the slab's bit arrays, group prefixes and the select table with bulk
// copies; its warps take groups of 32 units dynamically.  No big shared buffers, so the SM holds many more warps than
// in a fused kernel: the emission is a chain of shuffles and gathers and needs them to hide its latency.
// ------------------------------------------------------------------------------------------------------------------
struct EmitMisc { uint64_t bar; uint32_t next; uint32_t pad; };

template <int RB>
__global__ void __launch_bounds__(kEmitWarps * 32, VP_EMIT_MINB)
k_splat_emit(const VpWorldDev w, const uint32_t *__restrict__ ids, uint32_t n, const uint8_t *__restrict__ scratch, size_t arrived_bytes,
             uint8_t *__restrict__ arena)
{
	using G = Geo<RB>;
	constexpr int R = G::R, ZS = G::ZS, CL = G::CL;
	extern __shared__ __align__(128) uint8_t smem[];
	uint64_t *lv = reinterpret_cast<uint64_t *>(smem);
	uint8_t *lut = smem + G::E_OFF_LUT;
	uint32_t *gpre = reinterpret_cast<uint32_t *>(smem + G::E_OFF_GP);
	const uint32_t *xbs = gpre + G::XB_OFF;
	EmitMisc *misc = reinterpret_cast<EmitMisc *>(smem + G::E_OFF_MISC);

	const int tid = threadIdx.x, lane = tid & 31;
	uint4 *stage = reinterpret_cast<uint4 *>(smem + G::E_OFF_STAGE) + (tid >> 5) * 64;
	const uint32_t slab = gridDim.x - 1 - blockIdx.x;        // last written first: the tail of the scratch is still in L2
	const uint32_t chunk_i = slab / CL;
	const int crank = CL > 1 ? (int)(slab % CL) : 0;
	const SplatScratch sc = carve_scratch<RB>(const_cast<uint8_t *>(scratch), arrived_bytes, n);
	const uint32_t *rec = sc.rec + (size_t)slab * kSlabRec;
	const uint32_t c0 = rec[0], c1 = rec[1], c2 = rec[2], c3 = rec[3], c4 = rec[4];
	if ((c0 | c1 | c2 | c3 | c4) == 0) return;
	const unsigned long long *chrec = sc.chrec + (size_t)chunk_i * kChunkRec;
	const unsigned long long choff = chrec[0];
	if (choff == ~0ull) return;                               // arena overflow: nothing was reserved
	if (tid == 0) {
		mbar_init(&misc->bar, 1);
		mbar_fence_init();
		misc->next = 0;
		mbar_arrive_expect_tx(&misc->bar, G::LV_STRIDE * 8 + 2048 + G::GP_STRIDE * 4);
		tma_load_1d(lv, sc.pyr + (size_t)slab * G::LV_STRIDE, G::LV_STRIDE * 8, &misc->bar);
		tma_load_1d(lut, kSelLut.v, 2048, &misc->bar);
		tma_load_1d(gpre, sc.gp + (size_t)slab * G::GP_STRIDE, G::GP_STRIDE * 4, &misc->bar);
	}
	const uint32_t cid = ids[chunk_i];
	const uint32_t ccx = cid & ((1u << w.bits[0]) - 1), ccy = (cid >> w.bits[0]) & ((1u << w.bits[1]) - 1), ccz = cid >> (w.bits[0] + w.bits[1]);
	const int z0 = crank * ZS;
	const uint32_t b0 = rec[8], b1 = rec[9], b2 = rec[10], b3 = rec[11], b4 = rec[12];
	Ctx<RB> cx_{w, lv, reinterpret_cast<const uint8_t *>(chrec[1]), reinterpret_cast<const uint8_t *>(chrec[2]),
	            reinterpret_cast<const uint8_t *>(chrec[3]), reinterpret_cast<const uint8_t *>(chrec[4]), z0, ccx << RB, ccy << RB, ccz << RB,
	            reinterpret_cast<const uint8_t *>(chrec[1]) + (size_t)z0 * R * R, reinterpret_cast<const uint8_t *>(chrec[3]) + (size_t)z0 * R * R,
	            (uint32_t)z0 * R};
	uint2 *out2 = reinterpret_cast<uint2 *>(arena + choff);
	__syncthreads();
	if (tid < 32) mbar_wait(&misc->bar, 0);               // one warp polls the bulk copies, the others sleep in the barrier
	__syncthreads();

	// Slot s of a group belongs to the unit i with p_i <= s < p_i + c_i (shuffle binary search over the lanes'
	// exclusive prefixes) and inside the unit to its (s - p_i)-th set bit (popc select): every lane emits one splat
	// per round whatever the distribution of visible voxels, and the 8-byte stores of a warp are contiguous.
	const uint32_t nne = gpre[G::NG + 1];
	const uint8_t *glist = reinterpret_cast<const uint8_t *>(gpre + G::NG + 2);
#if VP_EMIT_STATIC
	for (uint32_t k = (uint32_t)(tid >> 5); k < nne; k += kEmitWarps) {
#else
	for (;;) {
		uint32_t k = 0;
		if (lane == 0) k = atomicAdd(&misc->next, 1u);
		k = __shfl_sync(0xffffffffu, k, 0);
		if (k >= nne) break;
#endif
		const int g = glist[k];
		const uint32_t gs = gpre[g];
		if (g < G::grp_off(1)) emit_group<RB, 0>(cx_, lv, xbs, lut, stage, g, out2 + b0 + (gs - gpre[0]), lane);
		else if (g < G::grp_off(2)) emit_group<RB, 1>(cx_, lv, xbs, lut, stage, g - G::grp_off(1), out2 + b1 + (gs - gpre[G::grp_off(1)]), lane);
		else if (g < G::grp_off(3)) emit_group<RB, 2>(cx_, lv, xbs, lut, stage, g - G::grp_off(2), out2 + b2 + (gs - gpre[G::grp_off(2)]), lane);
		else if (g < G::grp_off(4)) emit_group<RB, 3>(cx_, lv, xbs, lut, stage, g - G::grp_off(3), out2 + b3 + (gs - gpre[G::grp_off(3)]), lane);
		else emit_group<RB, 4>(cx_, lv, xbs, lut, stage, g - G::grp_off(4), out2 + b4 + (gs - gpre[G::grp_off(4)]), lane);
	}
}

// The per-chunk arrival counters sit at the start of the scratch, sized by the capacity it was allocated for.
inline size_t arrived_region_bytes(uint32_t cap_chunks) { return ((size_t)cap_chunks * 4 + 255) / 256 * 256; }

template <int RB>
cudaError_t launch(const VpWorldDev &w, const uint32_t *d_ids, uint32_t n, VpResultDev *d_results, const uint32_t *d_result_pos,
                   uint8_t *arena, VpArenaDev *state, uint8_t *scratch, uint32_t scratch_chunks, cudaStream_t s)
{
	using G = Geo<RB>;
	// the opt-in is per device (a process may hold contexts on several GPUs); the calls are cheap
	cudaError_t e = cudaFuncSetAttribute(k_splat_count<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(k_splat_emit<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::E_SMEM);
	if (e != cudaSuccess) return e;
	const size_t arrived_bytes = arrived_region_bytes(scratch_chunks);
	k_splat_count<RB><<<n * G::CL, G::THREADS, G::SMEM, s>>>(w, d_ids, n, scratch, arrived_bytes, d_results, d_result_pos, state);
	k_splat_emit<RB><<<n * G::CL, kEmitWarps * 32, G::E_SMEM, s>>>(w, d_ids, n, scratch, arrived_bytes, arena);
	return cudaGetLastError();
}

template <int RB> size_t scratch_bytes(uint32_t n)
{
	using G = Geo<RB>;
	const size_t slabs = (size_t)n * G::CL;
	return arrived_region_bytes(n) + slabs * ((size_t)G::LV_STRIDE * 8 + (size_t)G::GP_STRIDE * 4 + kSlabRec * 4) + (size_t)n * kChunkRec * 8 + 256;
}

} // namespace

cudaError_t vp_launch_splat(const VpWorldDev &w, const uint32_t *d_ids, uint32_t n, VpResultDev *d_results, const uint32_t *d_result_pos,
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

// Bytes of device scratch for splat rebuilds of up to n chunks per launch (arrival counters, which must be zero
// before the first launch, then bit arrays + prefixes + records of every slab).
size_t vp_splat_scratch_bytes(int rb, uint32_t n)
{
	switch (rb) { case 4: return scratch_bytes<4>(n); case 5: return scratch_bytes<5>(n); case 6: return scratch_bytes<6>(n); case 7: return scratch_bytes<7>(n); }
	return 0;
}

int vp_splat_smem_bytes(int rb)
{
	switch (rb) { case 4: return Geo<4>::SMEM; case 5: return Geo<5>::SMEM; case 6: return Geo<6>::SMEM; case 7: return Geo<7>::SMEM; }
	return -1;
}

#ifdef VP_PROFILE_PHASES
extern "C" __attribute__((visibility("default"))) int vp_debug_phase_cycles(unsigned long long out[8], int reset)
{
	cudaDeviceSynchronize();
	cudaError_t e = cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 8);
	if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z); }
	return (int)e;
}
#endif
