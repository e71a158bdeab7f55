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

#include "vp_worldgen_core.h"

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
		for (int k = 0; k < VPW_TREE_VOXELS; k++) {
			int32_t x, y, z; uint8_t v;
			if (tree_voxel(P, tx, ty, tz, k, &x, &y, &z, &v)) chunk_put(&ctx, x, y, z, v);
		}
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
