// vp_splat.cu -- cull + 5-level LOD + splat-list emission for a batch of chunks: ONE kernel, one pass over the voxels,
// no waiting between CTAs.
//
// Replaces, byte for byte, the splat branch of the reference dispatcher (chunkset.c:371-458):
//   chunk_make_mask (mesher.c:377-456) -> chunk_make_splatlist level 0 (mesher.c:497-536)
//   -> 4 x { chunk_mask_downsample (mesher.c:460-493) ; chunk_make_splatlist }
//
// k_splat -- one CTA per 16-slice z-slab of a chunk:
//   1. streams the slab (+ the slice below and above) through a ring of shared-memory tiles with 1-D TMA bulk copies
//      (cp.async.bulk + mbarrier), plus the +x / +y halo bytes of the neighbour chunks;
//   2. packs bytes into 64-bit occupancy rows (bit x of row (z,y)): SWAR non-zero flags gathered with IDP.4A;
//   3. derives the visibility rows with shift / AND-NOT face tests, then the 4 LOD levels by pair-OR-compress of the
//      bit rows -- no byte mask is ever materialised;
//   4. counts groups of 32 row words with popc + REDUX, one warp scans the group counts (stable z,y,x order comes from
//      prefixes, never from atomics);
//   5. per level: compacts the non-empty row words into an entry list {first slot, word index} + a table "round ->
//      first entry", then emits the splats in rounds of 32 consecutive slots: slot -> entry by popc over the mask of
//      entry starts inside the window, slot -> bit by popc halving + a table; position from the bit index, colour byte
//      gathered from the voxels this CTA streamed microseconds ago (L2), for LOD >= 1 the colour of the "last non-zero
//      child in scan order" found by descending the bit pyramids; shadow bit from two uint16 loads;
//   6. the records of a slab go to a staging arena (one atomicAdd per slab, its own counts only).  A chunk's buffer is
//      [L0|L1|L2|L3|L4] with the slabs in z order inside each level, so the final place of a record depends on the
//      counts of ALL slabs of the chunk: the LAST slab of a chunk to finish (arrival counter) reserves the chunk's
//      buffer with one atomicAdd and moves the (L2-resident) staged segments of all slabs to their place.  Nobody
//      waits for anybody, and the bit arrays never leave shared memory.
//
// The VP_* macros below are tuning hooks (scripts/build_variant.sh); the defaults are the measured best for 64^3 chunks.
#include "vp_device.cuh"
#include <cstddef>
using namespace vp;

namespace {

#ifndef VP_CW
#define VP_CW 4
#endif
#ifndef VP_MINB6
#define VP_MINB6 7
#endif
#ifndef VP_RING6
#define VP_RING6 4
#endif
#ifndef VP_PREFETCH
#define VP_PREFETCH 1
#endif
#ifndef VP_T6
#define VP_T6 256
#endif

template <int RB> struct Geo {
	static constexpr int R = 1 << RB;
	static constexpr int CW = RB == 6 ? VP_CW : 8;       // byte->bit consumer warps
	static constexpr int RING = RB == 6 ? VP_RING6 : 4;  // tiles in the TMA ring
	static constexpr int THREADS = RB == 6 ? VP_T6 : (CW + 1) * 32;
	static constexpr int NWARPS = THREADS / 32;
	static constexpr int PW = NWARPS - 1;                // producer warp (its lane 0 issues the bulk copies)
	static constexpr int MINB = RB == 6 ? VP_MINB6 : (RB < 6 ? 5 : 2);
	static constexpr int ZS = 16;                        // z slices per CTA
	static constexpr int CL = R / ZS;                    // CTAs (slabs) per chunk
	static constexpr int NW = R > 64 ? R / 64 : 1;       // 64-bit words per level-0 row
	static constexpr int SLICE = R * R;
	static constexpr int TILE = SLICE < 4096 ? SLICE : 4096;
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
	// "units" = the 64-bit words of the rows in emission order; groups of 32 units for counting
	__host__ __device__ static constexpr int nunits(int l) { return nrows(l) * NWl(l); }
	__host__ __device__ static constexpr int ngroups(int l) { return (nunits(l) + 31) / 32; }
	__host__ __device__ static constexpr int grp_off(int l) { return l == 0 ? 0 : grp_off(l - 1) + ngroups(l - 1); }
	static constexpr int NG = grp_off(5);
	// ---- shared memory carve-up (bytes).  [0, UNION_END) is used twice:
	//   streaming + visibility:  ring | halo | occ | occx      (the level bit arrays lv are built over the drained ring)
	//   emission:                lv | entry list | round table (occ / halo are dead by then)
	static constexpr int RING_BYTES = RING * TILE;
	static constexpr int LV_STRIDE = (LV_WORDS + 1) / 2 * 2;            // uint64 words, 16-byte multiple
	static constexpr int LV_BYTES = LV_STRIDE * 8;
	static constexpr int SCRATCH0 = ((RING_BYTES > LV_BYTES ? RING_BYTES : LV_BYTES) + 127) / 128 * 128;
	static constexpr int HALO_BYTES = 2 * ZS * R;
	static constexpr int OCC_WORDS = NSL * (R + 1) * NW;
	static constexpr int OCCX_WORDS = ZS * NW;
	static constexpr int OFF_HALO = SCRATCH0;
	static constexpr int OFF_OCC = OFF_HALO + HALO_BYTES;
	static constexpr int OFF_OCCX = OFF_OCC + OCC_WORDS * 8;
	static constexpr int COUNT_END = OFF_OCCX + OCCX_WORDS * 8;
	static constexpr int OFF_LV = 0;
	static constexpr int MAXU = nunits(0);                              // level 0 has the most units
	static constexpr int OFF_ENT = LV_BYTES;                            // uint2 per non-empty unit of the level being emitted
	static constexpr int TBL_N = MAXU * 65 / 32 + 2;                    // rounds of a level: a unit holds at most 65 splats
	static constexpr int OFF_TBL = OFF_ENT + MAXU * 8;
	static constexpr int EMIT_END = OFF_TBL + (TBL_N * 2 + 15) / 16 * 16;
	static constexpr int UNION_END = ((COUNT_END > EMIT_END ? COUNT_END : EMIT_END) + 127) / 128 * 128;
	static constexpr int OFF_LUT = UNION_END;
	static constexpr int OFF_BARS = OFF_LUT + 2048;
	static constexpr int OFF_MISC = OFF_BARS + (2 * RING + 2) * 8;
	static constexpr int MISC_BYTES = 128;
	static constexpr int SMEM = OFF_MISC + MISC_BYTES + (NG + 1) * 4 + 16;
	static_assert(MAXU <= 65535, "unit indices are stored in 16 bits");
	static constexpr int TOTU = nunits(0) + nunits(1) + nunits(2) + nunits(3) + nunits(4);
	static_assert(TOTU * 65 < (1 << 19) && TOTU < (1 << 13), "packed group prefix: 19 bits of slots, 13 bits of units");
	static_assert(5 * CL <= 64, "copy segments of a chunk");
};
constexpr int kSlabRec = 8;         // uint32 per slab record: [0..4] splats per level, [5] unused, [6,7] byte offset in the staging arena

struct Misc {                       // small per-CTA area in shared memory
	const uint8_t *src[4];          // own voxels / +x face plane / +y voxels / +z voxels of the chunk (null = null chunk)
	unsigned long long out_off;     // where this slab's records go: bytes into the staging arena (the chunk's arena when CL == 1), ~0 = no space
	unsigned long long chunk_off;   // copy phase: byte offset of the chunk's buffer in the arena
	uint32_t S[5];                  // splats per level of this slab
	uint32_t lbase[5];              // first record of level l inside the slab's staged records
	uint32_t total;
	uint32_t last;                  // this CTA was the last slab of its chunk to finish
	uint32_t ctot[5];               // copy phase: splats per level of the chunk
};
static_assert(sizeof(Misc) <= 128, "Misc must fit MISC_BYTES");

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

template <int RB> struct Ctx {
	using G = Geo<RB>;
	const VpWorldDev &w;
	const uint64_t *lv;             // level bit arrays (shared memory)
	const uint8_t *const *src;      // Misc::src (shared memory)
	int z0;
	uint32_t ox, oy, oz;            // chunk origin in world voxels

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

	// Voxel byte of the level-0 cell (X,Y,Z) of kind `kind` (Z relative to the slab): one byte gather from L2 (this CTA
	// streamed the voxel microseconds ago).  MAIN: own chunk; XPL: x = 0 column of the +x neighbour (its x-face plane
	// [z][y]); YPL: y = 0 row of the +y neighbour; ZPL: slice 0 of the +z neighbour.
	__device__ __forceinline__ uint32_t voxel(int kind, int X, int Y, int Z) const
	{
		const uint8_t *base = src[kind];
		const uint32_t zz = (uint32_t)(z0 + Z);
		uint32_t off;
		if (kind == MAIN) off = (((zz << RB) + (uint32_t)Y) << RB) + (uint32_t)X;
		else if (kind == XPL) off = (zz << RB) + (uint32_t)Y;
		else if (kind == YPL) off = (zz << (2 * RB)) + (uint32_t)X;
		else off = ((uint32_t)Y << RB) + (uint32_t)X;
		return __ldg(base + off);
	}

	// Colour byte of the level-L cell (X,Y,Z): the voxel found by descending the bit pyramids to the last non-zero child.
	template <int L>
	__device__ __forceinline__ uint32_t colour(int kind, int X, int Y, int Z) const
	{
		descend<L>(kind, X, Y, Z);
		return voxel(kind, X, Y, Z);
	}
};

// Unit u of level L: its 64-bit word and the +x plane bit that follows it in scan order (0/1).
template <int RB, int L>
__device__ __forceinline__ uint64_t load_unit(const uint64_t *lv, int u, uint32_t &xb)
{
	using G = Geo<RB>;
	constexpr int Rl = G::Rl(L), NWl = G::NWl(L), n_main = G::Zl(L) * (Rl + 1), U = G::nunits(L);
	xb = 0;
	if (u >= U) return 0ull;
	const uint64_t *base = lv + G::lvl_off(L);
	if (u < n_main * NWl) {
		if (u % NWl == NWl - 1) {
			const int q = u / NWl, Z = q / (Rl + 1), Y = q - Z * (Rl + 1);
			if (Y < Rl) xb = (uint32_t)(base[G::xpl_off(L) + Z * NWl + (Y >> 6)] >> (Y & 63)) & 1u;
		}
		return base[u];
	}
	return base[G::zpl_off(L) + (u - n_main * NWl)];
}

// The word alone (u < nunits(L)).
template <int RB, int L>
__device__ __forceinline__ uint64_t unit_word(const uint64_t *lv, uint32_t u)
{
	using G = Geo<RB>;
	constexpr uint32_t n_mw = G::main_words(L);                          // == n_main * NWl
	return lv[G::lvl_off(L) + u + (u >= n_mw ? (uint32_t)(G::zpl_off(L) - G::main_words(L)) : 0u)];
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

// Emission of level L of a slab by the whole CTA.
//   pass B: the non-empty units (row words) of the level are compacted, in scan order, into entries {first slot, unit};
//           tbl[r] = the entry that holds slot 32 r.
//   rounds: warp `warp` writes the slots [32 r, 32 r + 32) for r = warp, warp + NWARPS, ...: lane j reads entry tbl[r] + j,
//           a REDUX.OR builds the mask of entry starts inside the window, slot t belongs to entry popc(starts <= t); inside
//           the entry it is the (slot - first slot)-th set bit of the word (select64), or the +x plane cell that follows a
//           slab row in scan order (always the entry's last slot).  Every lane emits one splat per round whatever the
//           distribution of visible voxels, and the 8-byte stores of a warp are contiguous.
template <int RB, int L>
__device__ __forceinline__ void emit_level(const Ctx<RB> &cx, const uint64_t *lv, const uint8_t *lut, uint2 *ent, uint16_t *tbl,
                                           const uint32_t *gpre, uint32_t S, uint2 *out, int warp, int lane)
{
	using G = Geo<RB>;
	constexpr int R = G::R, Rl = G::Rl(L), NWl = G::NWl(L), n_main = G::Zl(L) * (Rl + 1), G0 = G::grp_off(L), NGl = G::ngroups(L);
	constexpr uint32_t FULL = 0xffffffffu;
	constexpr uint32_t XW = Rl < 64 ? Rl : 64;           // x of the +x plane cell relative to the unit's first bit
	constexpr uint32_t d = L ? (1u << L) : 0u;           // shadow sample offset of LOD splats (mesher.c:526-531)
	(void)R;
	const uint32_t pbase = gpre[G0], sbase = pbase & 0x7FFFFu, ebase = pbase >> 19;

	// ---- pass B -----------------------------------------------------------------------------------------
	for (int g = warp; g < NGl; g += G::NWARPS) {
		const uint32_t pg = gpre[G0 + g];
		if (((gpre[G0 + g + 1] - pg) & 0x7FFFFu) == 0u) continue;        // no splat in this group
		const int u = g * 32 + lane;
		uint32_t xb;
		const uint64_t word = load_unit<RB, L>(lv, u, xb);
		const uint32_t c = (uint32_t)__popcll(word) + xb;
		uint32_t inc = c;
		#pragma unroll
		for (int e = 1; e < 32; e <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, e); if (lane >= e) inc += t; }
		const uint32_t ne = __ballot_sync(FULL, c != 0u);
		if (c) {
			const uint32_t p = (pg & 0x7FFFFu) - sbase + inc - c;
			const uint32_t k = (pg >> 19) - ebase + (uint32_t)__popc(ne & ((1u << lane) - 1u));
			ent[k] = make_uint2(p, (uint32_t)u);
			for (uint32_t r = (p + 31u) >> 5; (r << 5) < p + c; r++) tbl[r] = (uint16_t)k;
		}
	}
	__syncthreads();

	// ---- rounds ------------------------------------------------------------------------------------------
	const uint32_t n_ent = (gpre[G0 + NGl] >> 19) - ebase;
	const uint32_t rounds = (S + 31u) >> 5;
	for (uint32_t r = (uint32_t)warp; r < rounds; r += G::NWARPS) {
		const uint32_t s0 = r << 5;
		const uint32_t j = (uint32_t)tbl[r] + (uint32_t)lane;
		uint2 e = make_uint2(0xFFFFFFFFu, 0u);
		if (j < n_ent) e = ent[j];
		// entries after the first start inside the window (the first one holds slot s0 itself)
		const uint32_t dd = e.x - s0;
		const uint32_t heads = __reduce_or_sync(FULL, (lane != 0 && dd < 32u) ? (1u << dd) : 0u);
		const uint32_t t = min((uint32_t)lane, S - 1u - s0);
		const uint32_t rel = (uint32_t)__popc(heads & (0xFFFFFFFFu >> (31u - t)));
		const uint32_t p = __shfl_sync(FULL, e.x, rel), u = __shfl_sync(FULL, e.y, rel);
		const uint64_t word = unit_word<RB, L>(lv, u);
		const uint32_t lo = (uint32_t)word, hi = (uint32_t)(word >> 32);
		const uint32_t k = s0 + t - p, wcl = __popc(lo), wcw = wcl + __popc(hi);
		const bool isx = k >= wcw;                       // the unit's +x plane cell (only ever its last slot)
		const uint32_t pos = isx ? XW : select64(lo, hi, wcl, k, lut);
		// cell coordinates at level L
		const uint32_t q = NWl == 1 ? u : u / NWl, x = (NWl == 1 ? 0u : (u % NWl) * 64u) + pos;
		const bool mainrow = q < (uint32_t)n_main;
		const uint32_t Zloc = mainrow ? q / (uint32_t)(Rl + 1) : 0u;
		const uint32_t Y = mainrow ? q - Zloc * (uint32_t)(Rl + 1) : q - (uint32_t)n_main;
		const uint32_t Zc = mainrow ? (uint32_t)(cx.z0 >> L) + Zloc : (uint32_t)Rl;
		const uint32_t wx = cx.ox + (x << L), wy = cx.oy + (Y << L), wz = cx.oz + (Zc << L);
		// shadow_sample (shadow.h:56-63): !(map[idx] < y+1 && map[idx+1] < y+1)
		const uint32_t lim = wy + d + 1u;
		const uint16_t *sp = cx.w.shadow + ((wx + d) + (wy + d) + cx.w.sh_w * (wz + d - cx.w.sh_z0));
		const uint32_t sa = __ldg(sp), sb = __ldg(sp + 1);        // both loads in flight together
		const int kind = !mainrow ? ZPL : (Y == (uint32_t)Rl ? YPL : (isx ? XPL : MAIN));
		const uint32_t col = cx.template colour<L>(kind, (int)x, (int)Y, (int)Zloc);
		const uint32_t sh = ((sa >= lim) | (sb >= lim)) ? 64u : 0u;
		if (s0 + (uint32_t)lane < S) {
			const uint32_t rl = (wx & 0xFFFFu) | (wy << 16);              // world size <= 32768 per axis: wy fits 16 bits
			const uint32_t rh = (wz & 0xFFFFu) | ((col | sh) << 16);
			out[s0 + lane] = make_uint2(rl, rh);
		}
	}
	__syncthreads();                                     // the next level reuses ent / tbl
}

// End of a slab (CL > 1).  Every thread of the CTA calls this.  The LAST slab of a chunk to get here (per-chunk arrival
// counter) adds up the level counts of all slabs, reserves the chunk's contiguous [L0|L1|L2|L3|L4] buffer in the arena
// with one atomicAdd, writes the result record (ChunkMD.svl_items[], chunkset.c:469-483) and moves the staged records of
// all slabs to their final place: level-major, slabs in z order inside a level.  Nobody waits.
template <int RB>
__device__ __forceinline__ void finish_chunk(uint8_t *smem, uint32_t chunk_i, uint32_t *arrived, const uint32_t *recs, VpResultDev *res,
                                             uint8_t *arena, VpArenaDev *st, const uint8_t *stage)
{
	using G = Geo<RB>;
	constexpr int CL = G::CL, NSEG = 5 * CL;
	Misc *misc = reinterpret_cast<Misc *>(smem + G::OFF_MISC);
	uint32_t *srec = reinterpret_cast<uint32_t *>(smem + G::OFF_ENT);                        // [CL][kSlabRec]
	unsigned long long *seg_src = reinterpret_cast<unsigned long long *>(smem + G::OFF_ENT + 512);    // [NSEG] staging byte offset
	unsigned long long *seg_dst = seg_src + 64;                                              // [NSEG] arena byte offset
	uint32_t *seg_n = reinterpret_cast<uint32_t *>(seg_dst + 64);                            // [NSEG] records
	static_assert(CL * kSlabRec * 4 <= 512 && 512 + 64 * 8 * 2 + 64 * 4 <= G::MAXU * 8, "copy tables fit the entry list area");
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

	__threadfence();                                   // this thread's staged records (and thread 0's slab record) are visible device-wide
	__syncthreads();
	if (tid == 0) {
		const uint32_t prev = atomicAdd(arrived + chunk_i, 1u);
		const uint32_t last = prev == (uint32_t)(CL - 1);
		if (last) arrived[chunk_i] = 0;                // ready for the next launch
		misc->last = last;
	}
	__syncthreads();
	if (!misc->last) return;
	__threadfence();
	if (tid < CL * kSlabRec) srec[tid] = __ldcg(recs + (size_t)chunk_i * CL * kSlabRec + tid);
	__syncthreads();
	if (tid == 0) {
		uint32_t total = 0;
		bool staged = true;
		for (int l = 0; l < 5; l++) {
			uint32_t t = 0;
			for (int r = 0; r < CL; r++) t += srec[r * kSlabRec + l];
			misc->ctot[l] = t; total += t;
		}
		for (int r = 0; r < CL; r++) if ((srec[r * kSlabRec + 6] & srec[r * kSlabRec + 7]) == 0xFFFFFFFFu) staged = false;
		unsigned long long off = 0;
		if (total) {
			const unsigned long long bytes = (unsigned long long)total * 8ull;
			off = atomicAdd(&st->cursor, bytes);
			if (off + bytes > st->capacity || !staged) { atomicExch(&st->overflow, 1u); off = ~0ull; }
		}
		misc->chunk_off = off; misc->total = total;
		res->svl_offset = off;
		res->svl_items_total = total * 4u;
		for (int l = 0; l < 5; l++) res->svl_items[l] = misc->ctot[l] * 4u;
	}
	__syncthreads();
	if (misc->total == 0 || misc->chunk_off == ~0ull) return;
	if (tid < NSEG) {
		const int l = tid / CL, r = tid % CL;
		uint32_t before_src = 0, before_dst = 0;
		for (int k = 0; k < l; k++) { before_src += srec[r * kSlabRec + k]; before_dst += misc->ctot[k]; }
		for (int k = 0; k < r; k++) before_dst += srec[k * kSlabRec + l];
		const unsigned long long so = (unsigned long long)srec[r * kSlabRec + 6] | ((unsigned long long)srec[r * kSlabRec + 7] << 32);
		seg_src[tid] = so + (unsigned long long)before_src * 8ull;
		seg_dst[tid] = misc->chunk_off + (unsigned long long)before_dst * 8ull;
		seg_n[tid] = srec[r * kSlabRec + l];
	}
	__syncthreads();
	for (int sg = warp; sg < NSEG; sg += G::NWARPS) {
		const uint32_t cnt = seg_n[sg];
		const uint2 *s = reinterpret_cast<const uint2 *>(stage + seg_src[sg]);
		uint2 *dd = reinterpret_cast<uint2 *>(arena + seg_dst[sg]);
		for (uint32_t i = (uint32_t)lane; i < cnt; i += 128u) {
			uint2 v0 = __ldcg(s + i), v1, v2, v3;
			const bool b1 = i + 32u < cnt, b2 = i + 64u < cnt, b3 = i + 96u < cnt;
			if (b1) v1 = __ldcg(s + i + 32);
			if (b2) v2 = __ldcg(s + i + 64);
			if (b3) v3 = __ldcg(s + i + 96);
			dd[i] = v0;
			if (b1) dd[i + 32] = v1;
			if (b2) dd[i + 64] = v2;
			if (b3) dd[i + 96] = v3;
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// The kernel: one CTA per 16-slice slab of a chunk.
// ------------------------------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(Geo<RB>::THREADS, Geo<RB>::MINB)
k_splat(const VpWorldDev w, const uint32_t *__restrict__ ids, uint32_t n, uint32_t *__restrict__ arrived, uint32_t *__restrict__ recs,
        VpResultDev *__restrict__ results, const uint32_t *__restrict__ result_pos, uint8_t *__restrict__ arena, VpArenaDev *__restrict__ st,
        uint8_t *__restrict__ stage, VpArenaDev *__restrict__ stage_st)
{
	using G = Geo<RB>;
	constexpr int R = G::R, ZS = G::ZS, CL = G::CL, NW = G::NW, TILE = G::TILE, TPS = G::TPS, NT = G::NT;
	constexpr int kConsumerWarps = G::CW, kThreads = G::THREADS, kRing = G::RING, kWarps = G::NWARPS;
	static_assert(kConsumerWarps % kRing == 0 && TILE % (kConsumerWarps / kRing) == 0, "consumer warps must tile the ring slots");
	static_assert(kConsumerWarps < kWarps, "one warp is the producer");
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t *ring = smem;
	uint8_t *halo = smem + G::OFF_HALO;
	uint64_t *occ = reinterpret_cast<uint64_t *>(smem + G::OFF_OCC);
	uint64_t *occx = reinterpret_cast<uint64_t *>(smem + G::OFF_OCCX);
	uint64_t *lv = reinterpret_cast<uint64_t *>(smem + G::OFF_LV);
	uint2 *ent = reinterpret_cast<uint2 *>(smem + G::OFF_ENT);
	uint16_t *tbl = reinterpret_cast<uint16_t *>(smem + G::OFF_TBL);
	uint8_t *lut = smem + G::OFF_LUT;
	uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + G::OFF_BARS);
	uint64_t *bar_empty = bar_full + kRing;
	uint64_t *bar_halo = bar_empty + kRing;
	uint64_t *bar_lut = bar_halo + 1;
	Misc *misc = reinterpret_cast<Misc *>(smem + G::OFF_MISC);
	uint32_t *gpre = reinterpret_cast<uint32_t *>(smem + G::OFF_MISC + G::MISC_BYTES);     // [NG + 1] packed exclusive prefix

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int crank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
	const uint32_t chunk_i = blockIdx.x / CL;
	const int z0 = crank * ZS;
	const bool top = (z0 + ZS == R);
	uint32_t *rec = recs + (size_t)blockIdx.x * kSlabRec;
	VpResultDev *res = results + (result_pos ? result_pos[chunk_i] : chunk_i);
	(void)n;

	// ---- phase 0: source pointers (4 threads, one slot lookup each), barriers, zeroed bit arrays -------
	const uint32_t cid = ids[chunk_i];
	const int ccx = (int)(cid & ((1u << w.bits[0]) - 1)), ccy = (int)((cid >> w.bits[0]) & ((1u << w.bits[1]) - 1));
	const int ccz = (int)(cid >> (w.bits[0] + w.bits[1]));
	if (tid < 4) {
		const int sl = chunk_slot(w, ccx + (tid == 1), ccy + (tid == 2), ccz + (tid == 3));
		const uint8_t *ptr = nullptr;
		if (sl >= 0) ptr = tid == 1 ? w.xlo_pool + (size_t)sl * R * R : w.vox_pool + (size_t)sl * R * R * R;
		misc->src[tid] = ptr;
	}
	if (tid == G::PW * 32) {
		for (int i = 0; i < kRing; i++) { mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, kConsumerWarps / kRing); }
		mbar_init(bar_halo, 1);
		mbar_init(bar_lut, 1);
		mbar_fence_init();
	}
	for (int i = tid; i < G::OCC_WORDS + G::OCCX_WORDS; i += kThreads) occ[i] = 0;      // occ, occx are contiguous
	__syncthreads();
	const uint8_t *own = misc->src[0], *nbx_xlo = misc->src[1], *nby = misc->src[2], *nbz = misc->src[3];

	// slab without anything visible: zero record, then the chunk bookkeeping (CL == 1: the memset result record is final)
	auto finish_empty = [&]() {
		if constexpr (CL > 1) {
			if (tid < kSlabRec) rec[tid] = 0;
			finish_chunk<RB>(smem, chunk_i, arrived, recs, res, arena, st, stage);
		}
	};
	if (!own && !nbx_xlo && !nby && !nbz) {           // mesher.c:404-409: nothing can be visible
		finish_empty();
		return;
	}

	// source of streamed slice s (z = z0 - 1 + s); nullptr = air or not needed
	auto slice_src = [&](int s) -> const uint8_t * {
		int z = z0 - 1 + s;
		if (z < 0) return nullptr;                     // no -z test at z = 0 (pair walk starts at A, mesher.c:421)
		if (z < R) return own ? own + (size_t)z * R * R : nullptr;
		return nbz;                                    // z == R: slice 0 of the +z neighbour
	};
	const bool have_halo = nbx_xlo || nby;

	// ---- phase 1: TMA producer (last warp) / byte->bit consumers (warps 0..CW-1) -------------------
	uint32_t any_solid = 0;
	if (warp == G::PW) {
		if (lane == 0) {
			mbar_arrive_expect_tx(bar_lut, 2048);          // select table of the emission (every exit below waits for it)
			tma_load_1d(lut, kSelLut.v, 2048, bar_lut);
#if VP_PREFETCH
			for (int sl = kRing / TPS; sl < G::NSL; sl++) {            // slices beyond the first ring fill
				const uint8_t *src = slice_src(sl);
				if (src) l2_prefetch(src, R * R);
			}
#endif
			if (have_halo) {
				mbar_arrive_expect_tx(bar_halo, (nbx_xlo ? ZS * R : 0) + (nby ? ZS * R : 0));
				if (nbx_xlo) tma_load_1d(halo, nbx_xlo + (size_t)z0 * R, ZS * R, bar_halo);
				if (nby) for (int z = 0; z < ZS; z++) tma_load_1d(halo + ZS * R + z * R, nby + (size_t)(z0 + z) * R * R, R, bar_halo);
			}
			for (int t = 0; t < NT; t++) {
				const int b = t % kRing, u = t / kRing;
				if (u > 0) mbar_wait(bar_empty + b, (u - 1) & 1);
				const uint8_t *src = slice_src(t / TPS);
				if (src) {
					mbar_arrive_expect_tx(bar_full + b, TILE);
					tma_load_1d(ring + b * TILE, src + (size_t)(t % TPS) * TILE, TILE, bar_full + b);
				} else {
					mbar_arrive(bar_full + b);          // keep the phases aligned, nothing to load
				}
			}
		}
	} else if (warp < kConsumerWarps) {
		// Each ring slot is drained by a FIXED set of warps (an equal part of the tile each), so a warp meets the
		// phases of its slot strictly in order -- with more consumers than slots a warp could otherwise run two
		// phases ahead and alias the mbarrier parity.
		constexpr int WPS = kConsumerWarps / kRing, PART = TILE / WPS;
		const int b = warp % kRing, hpart = warp / kRing;
		for (int t = b; t < NT; t += kRing) {
			const int u = t / kRing, s = t / TPS, part = t % TPS;
			mbar_wait(bar_full + b, u & 1);
			if (slice_src(s)) {
				const uint8_t *tb = ring + b * TILE;
				uint64_t *orow = occ + (size_t)(s * (R + 1) + part * G::RPT) * NW;
				if constexpr (PART >= 1024 && R >= 64) {
					// 32 bytes per lane and iteration: two conflict-free 16-byte reads 512 bytes apart.  Lane pairs
					// exchange their 16-bit masks with one shuffle; the even lane stores the 32-bit word of the first
					// read, the odd lane the word of the second, so every lane stores once.
					const uint32_t psel = (lane & 1) ? 0x3276u : 0x5410u;
					uint32_t *o32 = reinterpret_cast<uint32_t *>(orow);
					#pragma unroll 2
					for (int off = hpart * PART + lane * 16; off < (hpart + 1) * PART; off += 1024) {
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
					for (int off = hpart * PART + lane * 16; off < (hpart + 1) * PART; off += 512) {
						const uint4 q4 = *reinterpret_cast<const uint4 *>(tb + off);
						if (PART >= 512 && !__any_sync(0xffffffffu, (q4.x | q4.y | q4.z | q4.w) != 0u)) continue;
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
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_empty + b);
		}
		// halo rows: 0..ZS-1 = x = 0 column of the +x neighbour (bits over y), ZS..2ZS-1 = y = 0 row of +y
		if (have_halo) {
			mbar_wait(bar_halo, 0);
			for (int f = tid; f < 2 * ZS * G::LPR; f += kConsumerWarps * 32) {
				const int hr = f / G::LPR, bo = (f % G::LPR) * 16;
				const bool isx = hr < ZS;
				if (isx ? (nbx_xlo != nullptr) : (nby != nullptr)) {
					uint32_t m = nz16(*reinterpret_cast<const uint4 *>(halo + hr * R + bo));
					any_solid |= m;
					uint64_t *dst = isx ? occx + hr * NW : occ + (size_t)((hr - ZS + 1) * (R + 1) + R) * NW;
					if (R >= 32) {
						uint32_t v = m << (bo & 16);
						v |= __shfl_xor_sync(0xffffffffu, v, 1);
						if (!(lane & 1)) reinterpret_cast<uint32_t *>(dst)[bo >> 5] = v;
					} else {
						reinterpret_cast<uint32_t *>(dst)[0] = m;
					}
				}
			}
		}
	}
	// A slab without any solid voxel (own, slice above, +x/+y halo) has nothing visible.
	const bool nonempty = __syncthreads_or(any_solid != 0u) != 0;
	if (!nonempty) {
		mbar_wait(bar_lut, 0);                         // no bulk copy may be in flight into this CTA's shared memory at exit
		finish_empty();
		return;
	}
	uint64_t *lv0 = lv;
	for (int i = tid; i < G::LV_STRIDE; i += kThreads) lv[i] = 0;
	__syncthreads();

	// ---- phase 2: visibility rows (closed form of the pair walk, mesher.c:421-448) -------------------
	for (int f0 = warp * 32; f0 < ZS * R; f0 += kWarps * 32) {
		const int f = f0 + lane, zi = f >> RB, y = f & (R - 1), s = zi + 1;
		const uint64_t *o = occ + (size_t)(s * (R + 1) + y) * NW;
		const uint64_t xbit = (occx[zi * NW + (y >> 6)] >> (y & 63)) & 1ull;
		// 32 rows of air with no solid +x halo cell: nothing of them is visible and lv0 is pre-zeroed
		{
			uint64_t any = xbit;
			#pragma unroll
			for (int k = 0; k < NW; k++) any |= o[k];
			if (!__any_sync(0xffffffffu, any != 0ull)) continue;
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
		uint64_t *xp = lv0 + G::xpl_off(0);
		if (R >= 32) { if (lane == 0) reinterpret_cast<uint32_t *>(xp + zi * NW)[y >> 5] = bal; }
		else if ((lane & 15) == 0) xp[zi] = (bal >> lane) & 0xFFFFu;
	}
	for (int f = tid; f < (ZS + (top && nbz ? R : 0)) * NW; f += kThreads) {
		const int r = f / NW, k = f % NW;
		if (r < ZS) {            // +y plane row of slice r: solid in the neighbour, air below it in this chunk
			const uint64_t *o = occ + (size_t)((r + 1) * (R + 1)) * NW;
			lv0[(r * (R + 1) + R) * NW + k] = o[R * NW + k] & ~o[(R - 1) * NW + k];
		} else {                 // +z plane row y = r - ZS
			const int y = r - ZS;
			lv0[G::zpl_off(0) + y * NW + k] = occ[(size_t)((ZS + 1) * (R + 1) + y) * NW + k] & ~occ[(size_t)(ZS * (R + 1) + y) * NW + k];
		}
	}
	__syncthreads();

	// ---- phase 3: LOD pyramids, level l from l-1 by OR of the child rows + pair-OR-compress ---------
	#pragma unroll
	for (int l = 1; l < 5; l++) {
		const int c = l - 1;
		const int Rl = G::Rl(l), Zl = G::Zl(l), Rc = G::Rl(c), NWc = G::NWl(c);
		const uint64_t *cm = lv + G::lvl_off(c);
		uint64_t *pm = lv + G::lvl_off(l);
		const int n_main = Zl * (Rl + 1), n_all = n_main + Zl + Rl;
		for (int g = tid; g < n_all; g += kThreads) {
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
		}
		__syncthreads();
	}

	// ---- phase 4: counts.  The bit rows of all levels are cut into groups of 32 "units" (one 64-bit word
	// plus, for the last word of a slab row, the +x plane bit that follows it in scan order).  One warp
	// per group: popc + REDUX gives the group's splat count, a ballot its number of non-empty units; one
	// warp then scans the packed pairs.  Stable (z,y,x) order follows from the prefix, not from atomics. ---
	for (int g = warp; g < G::NG; g += kWarps) {
		uint32_t xb, c;
		if (g < G::grp_off(1)) { c = __popcll(load_unit<RB, 0>(lv, g * 32 + lane, xb)) + xb; }
		else if (g < G::grp_off(2)) { c = __popcll(load_unit<RB, 1>(lv, (g - G::grp_off(1)) * 32 + lane, xb)) + xb; }
		else if (g < G::grp_off(3)) { c = __popcll(load_unit<RB, 2>(lv, (g - G::grp_off(2)) * 32 + lane, xb)) + xb; }
		else if (g < G::grp_off(4)) { c = __popcll(load_unit<RB, 3>(lv, (g - G::grp_off(3)) * 32 + lane, xb)) + xb; }
		else { c = __popcll(load_unit<RB, 4>(lv, (g - G::grp_off(4)) * 32 + lane, xb)) + xb; }
		const uint32_t ne = __popc(__ballot_sync(0xffffffffu, c != 0u));
		c = __reduce_add_sync(0xffffffffu, c);
		if (lane == 0) gpre[g] = c | (ne << 19);
	}
	__syncthreads();
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
		uint32_t Sl = 0;
		if (lane < 5) Sl = (gpre[G::grp_off(lane + 1)] - gpre[G::grp_off(lane)]) & 0x7FFFFu;      // grp_off(5) == NG
		uint32_t lb = Sl, t;
		t = __shfl_up_sync(0xffffffffu, lb, 1); if (lane >= 1) lb += t;
		t = __shfl_up_sync(0xffffffffu, lb, 2); if (lane >= 2) lb += t;
		t = __shfl_up_sync(0xffffffffu, lb, 4); if (lane >= 4) lb += t;
		const uint32_t total = __shfl_sync(0xffffffffu, lb, 4);
		if (lane < 5) { misc->S[lane] = Sl; misc->lbase[lane] = lb - Sl; }
		// room for this slab's records: its own counts only -- the staging arena (CL > 1) or, when the slab is the whole
		// chunk, the chunk's final buffer
		unsigned long long off = 0;
		if (lane == 0) {
			VpArenaDev *a = CL > 1 ? stage_st : st;
			if (total) {
				const unsigned long long bytes = (unsigned long long)total * 8ull;
				off = atomicAdd(&a->cursor, bytes);
				if (off + bytes > a->capacity) { atomicExch(&st->overflow, 1u); off = ~0ull; }
			}
			misc->out_off = off; misc->total = total;
			if constexpr (CL > 1) {
				rec[5] = 0; rec[6] = (uint32_t)off; rec[7] = (uint32_t)(off >> 32);
			} else {
				res->svl_offset = off;
				res->svl_items_total = total * 4u;
			}
		}
		if constexpr (CL > 1) { if (lane < 5) rec[lane] = Sl; }
		else { if (lane < 5) res->svl_items[lane] = Sl * 4u; }
	}
	__syncthreads();

	// ---- phase 5: emission, level by level ------------------------------------------------------------
	if (misc->total != 0 && misc->out_off != ~0ull) {
		mbar_wait(bar_lut, 0);
		const Ctx<RB> cx{w, lv, misc->src, z0, (uint32_t)ccx << RB, (uint32_t)ccy << RB, (uint32_t)ccz << RB};
		uint2 *out = reinterpret_cast<uint2 *>((CL > 1 ? stage : arena) + misc->out_off);
		if (misc->S[0]) emit_level<RB, 0>(cx, lv, lut, ent, tbl, gpre, misc->S[0], out + misc->lbase[0], warp, lane);
		if (misc->S[1]) emit_level<RB, 1>(cx, lv, lut, ent, tbl, gpre, misc->S[1], out + misc->lbase[1], warp, lane);
		if (misc->S[2]) emit_level<RB, 2>(cx, lv, lut, ent, tbl, gpre, misc->S[2], out + misc->lbase[2], warp, lane);
		if (misc->S[3]) emit_level<RB, 3>(cx, lv, lut, ent, tbl, gpre, misc->S[3], out + misc->lbase[3], warp, lane);
		if (misc->S[4]) emit_level<RB, 4>(cx, lv, lut, ent, tbl, gpre, misc->S[4], out + misc->lbase[4], warp, lane);
	}
	if constexpr (CL > 1) finish_chunk<RB>(smem, chunk_i, arrived, recs, res, arena, st, stage);
}

// The per-chunk arrival counters sit at the start of the scratch, sized by the capacity it was allocated for.
inline size_t arrived_region_bytes(uint32_t cap_chunks) { return ((size_t)cap_chunks * 4 + 255) / 256 * 256; }

template <int RB>
cudaError_t launch(const VpWorldDev &w, const uint32_t *d_ids, uint32_t n, VpResultDev *d_results, const uint32_t *d_result_pos,
                   uint8_t *arena, VpArenaDev *state, uint8_t *stage, VpArenaDev *stage_state, uint8_t *scratch, uint32_t scratch_chunks,
                   cudaStream_t s)
{
	using G = Geo<RB>;
	// the opt-in is per device (a process may hold contexts on several GPUs); the call is cheap
	cudaError_t e = cudaFuncSetAttribute(k_splat<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
	if (e != cudaSuccess) return e;
	uint32_t *arrived = reinterpret_cast<uint32_t *>(scratch);
	uint32_t *recs = reinterpret_cast<uint32_t *>(scratch + arrived_region_bytes(scratch_chunks));
	k_splat<RB><<<n * G::CL, G::THREADS, G::SMEM, s>>>(w, d_ids, n, arrived, recs, d_results, d_result_pos, arena, state, stage, stage_state);
	return cudaGetLastError();
}

template <int RB> size_t scratch_bytes(uint32_t n)
{
	using G = Geo<RB>;
	return arrived_region_bytes(n) + (size_t)n * G::CL * kSlabRec * 4 + 256;
}

} // namespace

cudaError_t vp_launch_splat(const VpWorldDev &w, const uint32_t *d_ids, uint32_t n, VpResultDev *d_results, const uint32_t *d_result_pos,
                            uint8_t *arena, VpArenaDev *state, uint8_t *stage, VpArenaDev *stage_state, uint8_t *scratch,
                            uint32_t scratch_chunks, cudaStream_t s)
{
	if (n == 0) return cudaSuccess;
	if (n > scratch_chunks) return cudaErrorInvalidValue;
	switch (w.rb) {
	case 4: return launch<4>(w, d_ids, n, d_results, d_result_pos, arena, state, stage, stage_state, scratch, scratch_chunks, s);
	case 5: return launch<5>(w, d_ids, n, d_results, d_result_pos, arena, state, stage, stage_state, scratch, scratch_chunks, s);
	case 6: return launch<6>(w, d_ids, n, d_results, d_result_pos, arena, state, stage, stage_state, scratch, scratch_chunks, s);
	case 7: return launch<7>(w, d_ids, n, d_results, d_result_pos, arena, state, stage, stage_state, scratch, scratch_chunks, s);
	default: return cudaErrorInvalidValue;
	}
}

// Bytes of device scratch for splat rebuilds of up to n chunks per launch: arrival counters (which must be zero before
// the first launch; the kernel leaves them zero) and one small record per slab.
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
