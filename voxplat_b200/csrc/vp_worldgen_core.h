/*
 * vp_worldgen_core.h -- the pure functions of the deterministic integer world generator, shared by the host generator
 * (vp_worldgen.c, gcc) and the device generator (vp_worldgen_dev.cu, nvcc): one source, so the two cannot drift.
 * See vp_worldgen.c for what the generator is (and is not) relative to the reference's chunkset/gen.c.
 */
#ifndef VP_WORLDGEN_CORE_H
#define VP_WORLDGEN_CORE_H
#include <stdint.h>

#ifdef __CUDACC__
#define VPW_FN __host__ __device__ static inline
#else
#define VPW_FN static inline
#endif

typedef struct {
	uint32_t seed;
	int32_t  root_bitw;
	int32_t  bits[3];       /* chunk-count bit widths per axis (ChunkSet.max_bitw) */
} vpw_params;

VPW_FN uint32_t mix32(uint32_t h)
{
	h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
	return h;
}
VPW_FN uint32_t hash3(uint32_t seed, uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t h = seed * 0x9e3779b1u;
	h = mix32(h ^ (a * 0x85ebca77u + 0x165667b1u));
	h = mix32(h ^ (b * 0xc2b2ae3du + 0x27d4eb2fu));
	h = mix32(h ^ (c * 0x27d4eb2fu + 0x9e3779b1u));
	return h;
}

/* Smooth value noise in Q16: lattice period `p` voxels (power of two), channel `ch`. */
VPW_FN uint32_t vnoise_q16(uint32_t seed, uint32_t ch, uint32_t x, uint32_t z, uint32_t p_bits)
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
VPW_FN uint32_t ridge_q16(uint32_t n) { int32_t d = (int32_t)(2 * n) - 65536; if (d < 0) d = -d; return 65536u - (uint32_t)(d > 65536 ? 65536 : d); }

/* Column height in voxels, >= 8 (water level), following the shape of gen.c:89-142. */
VPW_FN int32_t column_height(const vpw_params *P, uint32_t x, uint32_t z)
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

VPW_FN uint8_t surface_colour(const vpw_params *P, uint32_t x, uint32_t z, int32_t h)
{
	if (h < 9) return 23;
	if (100 + (int32_t)(vnoise_q16(P->seed, 7, x, z, 2) >> 11) > h) return 8;       /* grass line ~100..132 */
	if (150 + (int32_t)(vnoise_q16(P->seed, 8, x, z, 1) >> 10) > h) return 42;
	return 63;
}

/* Tree test for grid cell (gx,gz) of the 10-voxel lattice; returns 1 and the trunk column if planted. */
VPW_FN int tree_at(const vpw_params *P, int32_t gx, int32_t gz, int32_t *tx, int32_t *tz, int32_t *ty)
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

/* Voxel k (0 .. VPW_TREE_VOXELS-1) of the tree rooted at (tx,ty,tz), in the order the generator writes them
 * (later writes win): gen.c:147-184 -- three canopy tiers 7/5/3 wide and 2 high with half the leaves dropped at random
 * (k < VPW_TREE_CANOPY), then 14 trunk voxels, then 2 leaf voxels on top.  Returns 0 for a dropped leaf. */
#define VPW_TREE_CANOPY 166
#define VPW_TREE_VOXELS 182
VPW_FN int tree_voxel(const vpw_params *P, int32_t tx, int32_t ty, int32_t tz, int k, int32_t *x, int32_t *y, int32_t *z, uint8_t *v)
{
	if (k >= VPW_TREE_CANOPY) {
		const int i = k - VPW_TREE_CANOPY;
		*x = tx; *y = ty + i; *z = tz; *v = i < 14 ? 36 : 4;
		return 1;
	}
	int tier = 0, kk = k;
	if (kk >= 98) { tier = 1; kk -= 98; if (kk >= 50) { tier = 2; kk -= 50; } }
	const int w = 7 - tier * 2;
	const int dx = kk / (2 * w), dy = (kk / w) % 2, dz = kk % w;            /* loops: x outer, y, z inner */
	const int32_t px = tx + 3 - tier - dx, py = ty + 5 + 4 * tier - dy, pz = tz + 3 - tier - dz;
	if (hash3(P->seed ^ 0x7ee5u, (uint32_t)px, (uint32_t)py, (uint32_t)pz) & 1) return 0;
	*x = px; *y = py; *z = pz; *v = 4;
	return 1;
}

#endif
