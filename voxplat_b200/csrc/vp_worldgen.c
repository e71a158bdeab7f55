/*
 * vp_worldgen.c -- deterministic, seeded, INTEGER-ONLY world generator (host C, OpenMP).
 *
 * Produces the fixed synthetic input that both the oracle and the CUDA path consume (SURVEY.md
 * section 8(d) "Inputs", 8(a') u8).  It is NOT a port of the reference's chunkset/gen.c: that file's
 * arithmetic lives in FastNoise (un-vendored, un-pinned) and is non-deterministic under OpenMP.  What is
 * kept is the *structure* of the reference world so the cull/mesh statistics are game-like:
 *   - a height field clamped at a water level of 8 with ridged hills      (gen.c:89-142)
 *   - per-column jitter of a few voxels                                   (gen.c:122-125)
 *   - fall-off towards the world edge                                     (gen.c:128-135)
 *   - body colour 21, surface colours 23 / 8 / 42 / 63                    (gen.c:305-323)
 *   - trees on a 10-voxel grid: 14-voxel trunk (36), 3-tier canopy (4)    (gen.c:147-184, :210-229)
 *   - shadow map filled by the shadow_place_update rule                   (shadow.h:77-89)
 * Everything is integer / fixed point, so the world is bit-identical on every machine and for any
 * thread count (pure function of (seed, x, y, z)); all values are < 64 like the reference palette.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VPW_EXPORT __attribute__((visibility("default")))

typedef struct {
	uint32_t seed;
	int32_t  root_bitw;
	int32_t  bits[3];       /* chunk-count bit widths per axis (ChunkSet.max_bitw) */
} vpw_params;

static inline uint32_t mix32(uint32_t h)
{
	h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
	return h;
}
static inline uint32_t hash3(uint32_t seed, uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t h = seed * 0x9e3779b1u;
	h = mix32(h ^ (a * 0x85ebca77u + 0x165667b1u));
	h = mix32(h ^ (b * 0xc2b2ae3du + 0x27d4eb2fu));
	h = mix32(h ^ (c * 0x27d4eb2fu + 0x9e3779b1u));
	return h;
}

/* Smooth value noise in Q16: lattice period `p` voxels (power of two), channel `ch`. */
static uint32_t vnoise_q16(uint32_t seed, uint32_t ch, uint32_t x, uint32_t z, uint32_t p_bits)
{
	uint32_t ix = x >> p_bits, iz = z >> p_bits;
	uint32_t fx = ((x - (ix << p_bits)) << 16) >> p_bits;        /* Q16 fraction */
	uint32_t fz = ((z - (iz << p_bits)) << 16) >> p_bits;
	/* smoothstep 3t^2-2t^3 in Q16 */
	uint64_t tx = fx, tz = fz;
	uint32_t sx = (uint32_t)((tx * tx * (3u * 65536u - 2u * tx)) >> 32);
	uint32_t sz = (uint32_t)((tz * tz * (3u * 65536u - 2u * tz)) >> 32);
	uint32_t v00 = hash3(seed, ch, ix, iz) & 0xFFFF, v10 = hash3(seed, ch, ix + 1, iz) & 0xFFFF;
	uint32_t v01 = hash3(seed, ch, ix, iz + 1) & 0xFFFF, v11 = hash3(seed, ch, ix + 1, iz + 1) & 0xFFFF;
	uint32_t a = v00 + (uint32_t)(((int64_t)((int32_t)v10 - (int32_t)v00) * sx) >> 16);
	uint32_t b = v01 + (uint32_t)(((int64_t)((int32_t)v11 - (int32_t)v01) * sx) >> 16);
	return a + (uint32_t)(((int64_t)((int32_t)b - (int32_t)a) * sz) >> 16);
}

/* ridge(n) = 1 - |2n-1| in Q16 */
static inline uint32_t ridge_q16(uint32_t n) { int32_t d = (int32_t)(2 * n) - 65536; if (d < 0) d = -d; return 65536u - (uint32_t)(d > 65536 ? 65536 : d); }

/* Column height in voxels, >= 8 (water level), following the shape of gen.c:89-142. */
static int32_t column_height(const vpw_params *P, uint32_t x, uint32_t z)
{
	uint32_t wx = 1u << (P->bits[0] + P->root_bitw), wz = 1u << (P->bits[2] + P->root_bitw);
	uint32_t wy = 1u << (P->bits[1] + P->root_bitw);
	uint32_t u = vnoise_q16(P->seed, 1, x, z, 8);                       /* continental mask, period 256 */
	uint32_t a0 = 65536u - (uint32_t)(((uint64_t)u * u) >> 16);
	uint32_t e0 = ridge_q16(vnoise_q16(P->seed, 2, x, z, 6));           /* period 64 */
	uint32_t e1 = (uint32_t)(((uint64_t)ridge_q16(vnoise_q16(P->seed, 3, x, z, 5)) * e0) >> 17);
	uint32_t e2 = (uint32_t)(((uint64_t)ridge_q16(vnoise_q16(P->seed, 4, x, z, 4)) * (e0 + e1)) >> 16) / 3;
	uint32_t e3 = (uint32_t)(((uint64_t)ridge_q16(vnoise_q16(P->seed, 5, x, z, 3)) * (e0 + e1 + e2)) >> 18);
	uint32_t s = (uint32_t)(((uint64_t)(e0 + e1 + e2 + e3) * a0) >> 16);     /* Q16, 0..~2.1 */
	/* ~ s^1.23 : blend of s and s^2 */
	uint32_t s2 = (uint32_t)(((uint64_t)s * s) >> 16);
	uint32_t t = (uint32_t)((50462ull * s + 15074ull * s2) >> 16);
	int32_t h = (int32_t)((t * 130u) >> 16);
	/* per-column jitter: 3*n(5x) + 2*n(10x) in the reference is effectively white noise in [-5,5] */
	uint32_t j = hash3(P->seed, 6, x, z);
	h += (int32_t)((j & 7) + ((j >> 3) & 3)) - 5 + (int32_t)((j >> 5) & 1);
	/* edge fall-off over the outer fifth of the world (gen.c:128-135), measured to the nearest x / z edge */
	int32_t c = (int32_t)(wx / 2), cz = (int32_t)(wz / 2);
	int32_t dx = (int32_t)x - c; if (dx < 0) dx = -dx; if (dx > c) dx = c;
	int32_t dz = (int32_t)z - cz; if (dz < 0) dz = -dz; if (dz > cz) dz = cz;
	int32_t ex = c - dx, ez = cz - dz, e = ex < ez ? ex : ez;               /* voxels to the nearest x / z edge */
	if (5 * e < c) h = (int32_t)(((int64_t)h * 5 * e) / c);
	h -= 60;                                                                 /* water level */
	if (h < 8) h = 8;
	if (h > (int32_t)wy - 24) h = (int32_t)wy - 24;                          /* leave room for a tree */
	if (h < 2) h = 2;
	return h;
}

static inline uint8_t surface_colour(const vpw_params *P, uint32_t x, uint32_t z, int32_t h)
{
	if (h < 9) return 23;
	if (100 + (int32_t)(vnoise_q16(P->seed, 7, x, z, 2) >> 11) > h) return 8;       /* grass line ~100..132 */
	if (150 + (int32_t)(vnoise_q16(P->seed, 8, x, z, 1) >> 10) > h) return 42;
	return 63;
}

/* Tree test for grid cell (gx,gz) of the 10-voxel lattice; returns 1 and the trunk column if planted. */
static int tree_at(const vpw_params *P, int32_t gx, int32_t gz, int32_t *tx, int32_t *tz, int32_t *ty)
{
	int32_t wx = 1 << (P->bits[0] + P->root_bitw), wz = 1 << (P->bits[2] + P->root_bitw);
	int32_t x = gx * 10 + 4, z = gz * 10 + 4;                   /* gen.c:211-212: for x=64; x<max-64; x+=10 */
	if (wx < 160 || wz < 160) { if (x < 8 || z < 8 || x >= wx - 8 || z >= wz - 8) return 0; }
	else if (x < 64 || z < 64 || x >= wx - 64 || z >= wz - 64) return 0;
	uint32_t r = hash3(P->seed, 9, (uint32_t)gx, (uint32_t)gz);
	uint32_t density = vnoise_q16(P->seed, 10, (uint32_t)x, (uint32_t)z, 5);      /* forest patches */
	if ((r & 0xFFFF) > density / 2) return 0;
	x += (int32_t)((r >> 16) & 7) - 3; z += (int32_t)((r >> 19) & 7) - 3;        /* +-5*noise jitter */
	int32_t h = column_height(P, (uint32_t)x, (uint32_t)z);
	if (!(h > 10 && h < 100)) return 0;                                        /* gen.c:223 */
	*tx = x; *tz = z; *ty = h;
	return 1;
}

/* Visit every voxel of the tree rooted at (tx,ty,tz): gen.c:147-184 (canopy tiers 7/5/3 wide, 2 high,
 * half the leaves dropped at random; 14 trunk voxels; 2 leaf voxels on top). */
typedef void (*tree_cb)(void *ctx, int32_t x, int32_t y, int32_t z, uint8_t v);
static void tree_visit(const vpw_params *P, int32_t tx, int32_t ty, int32_t tz, tree_cb cb, void *ctx)
{
	int32_t bx = tx + 3, by = ty + 5, bz = tz + 3;
	for (int i = 0; i < 3; i++) {
		int w = 7 - i * 2;
		for (int x = 0; x < w; x++) for (int y = 0; y < 2; y++) for (int z = 0; z < w; z++) {
			int32_t px = bx - x, py = by - y, pz = bz - z;
			if (hash3(P->seed ^ 0x7ee5u, (uint32_t)px, (uint32_t)py, (uint32_t)pz) & 1) continue;
			cb(ctx, px, py, pz, 4);
		}
		bx -= 1; by += 4; bz -= 1;
	}
	for (int i = 0; i < 14; i++) cb(ctx, tx, ty + i, tz, 36);
	cb(ctx, tx, ty + 14, tz, 4);
	cb(ctx, tx, ty + 15, tz, 4);
}

struct chunk_ctx { uint8_t *out; int32_t ox, oy, oz, R, rb; uint32_t written; };
static void chunk_put(void *vctx, int32_t x, int32_t y, int32_t z, uint8_t v)
{
	struct chunk_ctx *c = vctx;
	int32_t lx = x - c->ox, ly = y - c->oy, lz = z - c->oz;
	if (lx < 0 || ly < 0 || lz < 0 || lx >= c->R || ly >= c->R || lz >= c->R) return;
	if (y < 2) return;                                                        /* edit.c:151 */
	c->out[((lz << c->rb | ly) << c->rb) | lx] = v;
	c->written++;
}

VPW_EXPORT int32_t vpw_height(const vpw_params *P, uint32_t x, uint32_t z) { return column_height(P, x, z); }

/* Fill one dense chunk (R^3 bytes, x fastest: chunkset.h:171-188).  Returns the number of solid voxels. */
VPW_EXPORT uint32_t vpw_gen_chunk(const vpw_params *P, uint32_t cx, uint32_t cy, uint32_t cz, uint8_t *out)
{
	int32_t rb = P->root_bitw, R = 1 << rb;
	int32_t ox = (int32_t)cx << rb, oy = (int32_t)cy << rb, oz = (int32_t)cz << rb;
	memset(out, 0, (size_t)R * R * R);
	uint32_t solid = 0;
	for (int32_t lz = 0; lz < R; lz++) for (int32_t lx = 0; lx < R; lx++) {
		int32_t h = column_height(P, (uint32_t)(ox + lx), (uint32_t)(oz + lz));
		int32_t top = h - oy; if (top > R) top = R;
		for (int32_t ly = 0; ly < top; ly++) out[((lz << rb | ly) << rb) | lx] = 21;
		if (top > 0) solid += (uint32_t)top;
		int32_t ls = h - 1 - oy;
		if (ls >= 0 && ls < R) out[((lz << rb | ls) << rb) | lx] = surface_colour(P, (uint32_t)(ox + lx), (uint32_t)(oz + lz), h);
	}
	/* trees whose canopy (reach +-4) or trunk can touch this chunk */
	struct chunk_ctx ctx = { out, ox, oy, oz, R, rb, 0 };
	int32_t g0x = (ox - 16) / 10 - 1, g1x = (ox + R + 16) / 10 + 1, g0z = (oz - 16) / 10 - 1, g1z = (oz + R + 16) / 10 + 1;
	for (int32_t gz = g0z < 0 ? 0 : g0z; gz <= g1z; gz++) for (int32_t gx = g0x < 0 ? 0 : g0x; gx <= g1x; gx++) {
		int32_t tx, ty, tz;
		if (!tree_at(P, gx, gz, &tx, &tz, &ty)) continue;
		if (ty + 16 < oy || ty >= oy + R) continue;
		tree_visit(P, tx, ty, tz, chunk_put, &ctx);
	}
	if (ctx.written) { solid = 0; for (int32_t i = 0; i < R * R * R; i++) solid += out[i] != 0; }
	return solid;
}

/* Generate `n` chunks (linear ids, chunkset.c:124-126 order: x fastest, then y, then z) into one
 * contiguous dense buffer; solid[k] receives the solid-voxel count (0 => the chunk is all air). */
VPW_EXPORT void vpw_gen_chunks(const vpw_params *P, const uint32_t *ids, uint32_t n, uint8_t *out, uint32_t *solid)
{
	size_t N = (size_t)1 << (3 * P->root_bitw);
	#pragma omp parallel for schedule(dynamic, 4)
	for (uint32_t k = 0; k < n; k++) {
		uint32_t id = ids[k];
		uint32_t cx = id & ((1u << P->bits[0]) - 1), cy = (id >> P->bits[0]) & ((1u << P->bits[1]) - 1);
		uint32_t cz = id >> (P->bits[0] + P->bits[1]);
		uint32_t s = vpw_gen_chunk(P, cx, cy, cz, out + N * k);
		if (solid) solid[k] = s;
	}
}

/*
 * Shadow map rows [z0, z1) from ANY dense world, by the reference's placement rule applied in a FIXED
 * order (z rows independent; inside a row x ascending, then y ascending):
 *     idx = (x+y) + SH*z ;  if (map[idx] >= y+1 || map[idx+1] >= y+1) skip ; else map[idx] = y
 * (shadow.h:45-63,77-89).  `chunks[id]` points at the dense voxels of chunk `id` or is NULL for air.
 * `map` addresses row z0 (SH = X+Y entries per row); it must be zero-filled and have >= 2 entries of
 * slack after the last row.
 */
VPW_EXPORT void vpw_shadow_rows(const vpw_params *P, const uint8_t *const *chunks, uint32_t z0, uint32_t z1, uint16_t *map)
{
	int32_t rb = P->root_bitw, R = 1 << rb;
	uint32_t X = 1u << (P->bits[0] + rb), Y = 1u << (P->bits[1] + rb), SH = X + Y;
	#pragma omp parallel for schedule(dynamic, 1)
	for (uint32_t z = z0; z < z1; z++) {
		uint16_t *row = map + (size_t)(z - z0) * SH;
		uint32_t cz = z >> rb, lz = z & (uint32_t)(R - 1);
		for (uint32_t x = 0; x < X; x++) {
			uint32_t cx = x >> rb, lx = x & (uint32_t)(R - 1);
			for (uint32_t cy = 0; cy < (1u << P->bits[1]); cy++) {
				const uint8_t *c = chunks[((cz << P->bits[1] | cy) << P->bits[0]) | cx];
				if (!c) continue;
				const uint8_t *col = c + ((size_t)lz << (2 * rb)) + lx;
				for (uint32_t ly = 0; ly < (uint32_t)R; ly++) {
					if (!col[(size_t)ly << rb]) continue;
					uint32_t y = (cy << rb) + ly, idx = x + y;
					if (row[idx] >= y + 1 || row[idx + 1] >= y + 1) continue;
					row[idx] = (uint16_t)y;
				}
			}
		}
	}
}
