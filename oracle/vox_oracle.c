/*
 * vox_oracle.c -- TEST INFRASTRUCTURE ONLY: a plain-C CPU restatement of the reference's chunk-rebuild
 * algorithm, used solely as the parity checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  The product (voxplat_b200/, include/) never includes, links or calls this file.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors for this path (SURVEY.md
 * section 4), so this restatement is pinned by differential execution against the compiled, unmodified
 * reference (oracle/_ref/libvoxref.so, built by oracle/Makefile from /root/reference) in
 * tests/test_oracle_vs_reference.py, and by the committed fixtures tests/golden/ that were generated
 * from that library (tests/golden/make_golden.py).
 *
 * It is a RESTATEMENT, not a copy: the reference walks voxel pairs and scatters into a byte mask; here
 * every function is written from the closed-form rules of SURVEY.md section 8(a) with a world-space
 * voxel getter.  Each function cites the reference lines whose behaviour it reproduces.
 *
 * World model: root R = 1<<rb voxels per chunk edge, chunk grid 2^bits[0] x 2^bits[1] x 2^bits[2];
 * chunk id = (cz<<by | cy)<<bx | cx (chunkset.c:124-126); voxel index = (z<<rb | y)<<rb | x
 * (chunkset.c:128-130).  chunks[id]==NULL means "all air" (the reference's null chunk).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define VO_EXPORT __attribute__((visibility("default")))

typedef struct {
	int32_t rb;
	int32_t bits[3];
	const uint8_t *const *chunks;
	const uint16_t *shadow;          /* (X+Y)*Z entries + >= 17*(X+Y)+64 zero padding (u3) */
} vo_world;

static inline uint32_t wdim(const vo_world *w, int a) { return 1u << (w->bits[a] + w->rb); }

/* chunkset_edit_read, edit.c:10-36: world-space fetch, 0 outside the world (negative coordinates
 * arrive as huge unsigned values and fail the bound test, edit.c:22-25). */
static inline uint8_t vox_at(const vo_world *w, uint32_t x, uint32_t y, uint32_t z)
{
	if (x >= wdim(w, 0) || y >= wdim(w, 1) || z >= wdim(w, 2)) return 0;
	uint32_t rb = (uint32_t)w->rb, m = (1u << rb) - 1;
	uint32_t id = (((z >> rb) << w->bits[1] | (y >> rb)) << w->bits[0]) | (x >> rb);
	const uint8_t *c = w->chunks[id];
	return c ? c[(((z & m) << rb | (y & m)) << rb) | (x & m)] : 0;
}
VO_EXPORT int vo_voxel(const vo_world *w, uint32_t x, uint32_t y, uint32_t z) { return vox_at(w, x, y, z); }

/* shadow_index / shadow_sample / shadow_sample_normal, shadow.h:45-75.  All arithmetic is uint32 like
 * the reference; when y+1 wraps to 0 the comparison "entry < 0" is false for any entry, so the sample
 * is 1 without touching the map (the reference's compiled code elides that dead load). */
static inline int shadow_pair(const vo_world *w, uint32_t x, uint32_t y, uint32_t z, int second)
{
	uint32_t lim = y + 1u;
	if (lim == 0) return 1;
	uint32_t sh = wdim(w, 0) + wdim(w, 1);
	uint32_t idx = (x + y) + sh * z;
	return !(w->shadow[idx] < lim && w->shadow[idx + second] < lim);
}
VO_EXPORT int vo_shadow(const vo_world *w, uint32_t x, uint32_t y, uint32_t z) { return shadow_pair(w, x, y, z, 1); }
VO_EXPORT int vo_shadow_normal(const vo_world *w, uint32_t x, uint32_t y, uint32_t z) { return shadow_pair(w, x, y, z, -1); }

static inline void chunk_origin(const vo_world *w, uint32_t id, uint32_t o[3])
{
	o[0] = (id & ((1u << w->bits[0]) - 1)) << w->rb;
	o[1] = ((id >> w->bits[0]) & ((1u << w->bits[1]) - 1)) << w->rb;
	o[2] = (id >> (w->bits[0] + w->bits[1])) << w->rb;
}

/*
 * Level-0 visibility mask, (R+1)^3 bytes with stride R+1 (chunk_make_mask, mesher.c:377-456, indexing
 * flatten1_no_po2 :365-374).  Closed form of the pair walk (:421-448):
 *   p in [0,R)^3      : mask[p] = v(p) iff v(p)!=0 and ( some p+e_i is air  or  some p-e_i with p_i>=1 is air )
 *   p with one p_i==R : mask[p] = v(p) iff v(p)!=0 and p-e_i is air          (v(p) lives in the +i neighbour)
 * where cells outside the world are air (null chunk, :393-394).  Faces at local coordinate 0 towards
 * the lower neighbour are NOT tested here; the lower neighbour emits that voxel at its index R.
 */
VO_EXPORT void vo_mask(const vo_world *w, uint32_t id, uint8_t *mask)
{
	uint32_t R = 1u << w->rb, S = R + 1, o[3];
	chunk_origin(w, id, o);
	memset(mask, 0, (size_t)S * S * S);
	for (uint32_t z = 0; z < R; z++) for (uint32_t y = 0; y < R; y++) for (uint32_t x = 0; x < R; x++) {
		uint8_t v = vox_at(w, o[0] + x, o[1] + y, o[2] + z);
		if (!v) continue;
		int open = !vox_at(w, o[0] + x + 1, o[1] + y, o[2] + z) || !vox_at(w, o[0] + x, o[1] + y + 1, o[2] + z)
		        || !vox_at(w, o[0] + x, o[1] + y, o[2] + z + 1)
		        || (x && !vox_at(w, o[0] + x - 1, o[1] + y, o[2] + z))
		        || (y && !vox_at(w, o[0] + x, o[1] + y - 1, o[2] + z))
		        || (z && !vox_at(w, o[0] + x, o[1] + y, o[2] + z - 1));
		if (open) mask[(z * S + y) * S + x] = v;
	}
	for (uint32_t a = 0; a < R; a++) for (uint32_t b = 0; b < R; b++) {
		uint8_t v;
		if ((v = vox_at(w, o[0] + R, o[1] + a, o[2] + b)) && !vox_at(w, o[0] + R - 1, o[1] + a, o[2] + b)) mask[(b * S + a) * S + R] = v;
		if ((v = vox_at(w, o[0] + a, o[1] + R, o[2] + b)) && !vox_at(w, o[0] + a, o[1] + R - 1, o[2] + b)) mask[(b * S + R) * S + a] = v;
		if ((v = vox_at(w, o[0] + a, o[1] + b, o[2] + R)) && !vox_at(w, o[0] + a, o[1] + b, o[2] + R - 1)) mask[(R * S + b) * S + a] = v;
	}
}

/*
 * LOD reduction of one level (chunk_mask_downsample, mesher.c:460-493): the source grid of level
 * `level-1` spans [0, (R>>(level-1))] per axis inside a stride-(R+1) buffer; a parent cell takes the
 * value of its LAST non-zero child in (z,y,x) scan order (:474-490).  Stated here as a gather: for each
 * parent, visit its (up to) 8 children in descending scan order and keep the first non-zero.
 * `dst` is fully rewritten (the reference relies on the caller's memset, chunkset.c:321,408,427,444).
 */
VO_EXPORT void vo_downsample(int32_t rb, int32_t level, const uint8_t *src, uint8_t *dst)
{
	uint32_t R = 1u << rb, S = R + 1;
	uint32_t src_ext = (R >> (level - 1)) + 1;            /* cells per axis in the source level */
	uint32_t dst_ext = ((src_ext - 1) >> 1) + 1;
	memset(dst, 0, (size_t)S * S * S);
	for (uint32_t Z = 0; Z < dst_ext; Z++) for (uint32_t Y = 0; Y < dst_ext; Y++) for (uint32_t X = 0; X < dst_ext; X++) {
		uint8_t v = 0;
		for (int k = 7; k >= 0 && !v; k--) {
			uint32_t cx = 2 * X + (k & 1), cy = 2 * Y + ((k >> 1) & 1), cz = 2 * Z + ((k >> 2) & 1);
			if (cx >= src_ext || cy >= src_ext || cz >= src_ext) continue;
			v = src[(cz * S + cy) * S + cx];
		}
		dst[(Z * S + Y) * S + X] = v;
	}
}

/*
 * Splat list of one level (chunk_make_splatlist, mesher.c:497-536): stable (z,y,x) compaction of the
 * non-zero mask cells; each emits int16 x,y,z = chunk origin + (cell << level) (truncating store, :523-524)
 * and int16 colour = mask | shadow<<6, the shadow sampled at the position + (1<<level) on all axes when
 * level>0 (:526-531).  Returns the number of int16 items written.
 */
VO_EXPORT uint32_t vo_splatlist(const vo_world *w, uint32_t id, int32_t level, const uint8_t *mask, int16_t *out)
{
	uint32_t R = 1u << w->rb, S = R + 1, ext = (R >> level) + 1, o[3], n = 0;
	chunk_origin(w, id, o);
	for (uint32_t z = 0; z < ext; z++) for (uint32_t y = 0; y < ext; y++) for (uint32_t x = 0; x < ext; x++) {
		uint8_t m = mask[(z * S + y) * S + x];
		if (!m) continue;
		uint32_t wx = o[0] + (x << level), wy = o[1] + (y << level), wz = o[2] + (z << level);
		out[n++] = (int16_t)wx; out[n++] = (int16_t)wy; out[n++] = (int16_t)wz;
		uint32_t d = level ? (1u << level) : 0;
		out[n++] = (int16_t)(m | (shadow_pair(w, wx + d, wy + d, wz + d, 1) << 6));
	}
	return n;
}

/* The dispatcher's splat sequence (chunkset.c:371-458): level 0 from the mask, levels 1..4 by repeated
 * reduction, the five segments concatenated; items[l] = int16 count of level l.  Returns the total. */
VO_EXPORT uint32_t vo_chunk_splat(const vo_world *w, uint32_t id, int16_t *out, uint32_t items[5])
{
	uint32_t R = 1u << w->rb, S = R + 1;
	size_t M = (size_t)S * S * S;
	uint8_t *a = malloc(M), *b = malloc(M), *t;
	uint32_t n = 0;
	vo_mask(w, id, a);
	for (int32_t l = 0; l < 5; l++) {
		if (l) { vo_downsample(w->rb, l, a, b); t = a; a = b; b = t; }
		items[l] = vo_splatlist(w, id, l, a, out + n);
		n += items[l];
	}
	free(a); free(b);
	return n;
}

/* Quad corner offsets per axis as bit triples (bit0=x, bit1=y, bit2=z): the first three blocks of the
 * reference's vertex table (mesher.c:19-33); the index patterns of mesher.c:53-65 for
 * [normal][rotated][6]. */
static const uint8_t quad_corner[3][4] = { {1, 5, 7, 3}, {2, 3, 7, 6}, {5, 4, 6, 7} };
static const uint8_t quad_index[2][2][6] = { { {0, 3, 1, 2, 1, 3}, {3, 2, 0, 1, 0, 2} },
                                             { {1, 3, 0, 3, 1, 2}, {0, 2, 3, 2, 0, 1} } };
/* which quad vertex receives AO term k, per axis (mesher.c:257-272) */
static const uint8_t ao_vertex[3][4] = { {0, 3, 2, 1}, {0, 1, 2, 3}, {1, 0, 3, 2} };

/*
 * Near-field quad mesh of one chunk (chunk_make_mesh, mesher.c:184-357).  For every voxel A of the
 * chunk in (z,y,x) order and every axis i in x,y,z: B = the voxel at A+e_i (neighbour chunk or air
 * outside the world, :200-211,:229-232).  If exactly one of A,B is solid, one quad is emitted:
 *   AIR / BLOCK cells (:239-244); 4 edge + 4 corner occupancy samples around AIR in the plane normal to
 *   i (sample_ao :72-116 == sample_ao_border :119-171 == world-space reads, 0 outside the world);
 *   per-vertex ao and its axis-dependent permutation (:256-272); index rotation when
 *   ao0+ao2 < ao1+ao3 (:274-276); shadow / diamond bits (:295-318); 4 vertices of int16 x,y,z,data
 *   (:321-339) and 6 uint32 indices (:345-349).
 * nv / ni are counts of int16 / uint32 ELEMENTS (chunkset.c:348,357).
 */
VO_EXPORT void vo_chunk_mesh(const vo_world *w, uint32_t id, int16_t *vbo, uint32_t *ibo, uint32_t *nv, uint32_t *ni)
{
	uint32_t R = 1u << w->rb, o[3], v = 0, k = 0, base = 0;
	chunk_origin(w, id, o);
	for (uint32_t z = 0; z < R; z++) for (uint32_t y = 0; y < R; y++) for (uint32_t x = 0; x < R; x++) {
		uint32_t p[3] = { o[0] + x, o[1] + y, o[2] + z };
		uint8_t A = vox_at(w, p[0], p[1], p[2]);
		for (int i = 0; i < 3; i++) {
			uint32_t q[3] = { p[0], p[1], p[2] };
			q[i]++;
			uint8_t B = vox_at(w, q[0], q[1], q[2]);
			if (!A == !B) continue;
			int normal = (A == 0);
			const uint32_t *air = B ? p : q, *blk = A ? p : q;
			int u0 = i == 0 ? 1 : 0, u1 = i == 2 ? 1 : 2;
			uint32_t s[3];
			int n[4], c[4];
			#define OCC(d0, d1) (memcpy(s, air, sizeof s), s[u0] += (uint32_t)(d0), s[u1] += (uint32_t)(d1), vox_at(w, s[0], s[1], s[2]) != 0)
			n[0] = OCC(-1, 0); n[1] = OCC(0, -1); n[2] = OCC(1, 0); n[3] = OCC(0, 1);
			c[0] = OCC(-1, -1); c[1] = OCC(1, -1); c[2] = OCC(1, 1); c[3] = OCC(-1, 1);
			#undef OCC
			int vao[4];
			for (int t = 0; t < 4; t++) vao[ao_vertex[i][t]] = (n[t] + n[(t + 1) & 3]) | c[t];
			int rotated = (vao[0] + vao[2] < vao[1] + vao[3]);
			uint32_t ws[3] = { blk[0], blk[1], blk[2] };
			int shadow = 0, diamond = 3;
			if (i == 0 && normal) shadow = shadow_pair(w, ws[0], ws[1], ws[2], -1);
			else if (i == 1 && !normal) shadow = shadow_pair(w, ws[0], ws[1], ws[2], 1);
			ws[i] -= (uint32_t)normal;
			if (i == 2) {
				shadow = 0;
				/* the two samples sit at the AIR cell and the cell below it (:306-317) */
				diamond = (!shadow_pair(w, air[0], air[1], air[2], 1)) << 1;
				diamond |= !shadow_pair(w, air[0], air[1] - 1u, air[2], 1);
			}
			for (int t = 0; t < 4; t++) {
				uint8_t cr = quad_corner[i][t];
				vbo[v++] = (int16_t)(ws[0] + (cr & 1));
				vbo[v++] = (int16_t)(ws[1] + ((cr >> 1) & 1));
				vbo[v++] = (int16_t)(ws[2] + ((cr >> 2) & 1));
				vbo[v++] = (int16_t)((A | B) | (vao[t] << 6) | ((i + 3 * normal) << 8) | (t << 11) | (shadow << 13) | (diamond << 14));
			}
			for (int t = 0; t < 6; t++) ibo[k++] = base + quad_index[normal][rotated][t];
			base += 4;
		}
	}
	*nv = v; *ni = k;
}

/* Count-only variant so callers can size the buffers exactly (the reference guesses, u7). */
VO_EXPORT uint32_t vo_chunk_mesh_faces(const vo_world *w, uint32_t id)
{
	uint32_t R = 1u << w->rb, o[3], f = 0;
	chunk_origin(w, id, o);
	for (uint32_t z = 0; z < R; z++) for (uint32_t y = 0; y < R; y++) for (uint32_t x = 0; x < R; x++) {
		int a = vox_at(w, o[0] + x, o[1] + y, o[2] + z) != 0;
		f += a != (vox_at(w, o[0] + x + 1, o[1] + y, o[2] + z) != 0);
		f += a != (vox_at(w, o[0] + x, o[1] + y + 1, o[2] + z) != 0);
		f += a != (vox_at(w, o[0] + x, o[1] + y, o[2] + z + 1) != 0);
	}
	return f;
}

/*
 * RLE codec (rle.c): 32-bit little-endian words, run length in the low 24 bits, value in the top 8,
 * a 0 word terminates.  Encode (rle_compress :44-87): maximal runs, a run is cut when its count
 * reaches 0xFFFFFF (:62).  Returns the word count including the terminator; `out` may be NULL.
 */
VO_EXPORT uint32_t vo_rle_encode(const uint8_t *v, uint32_t n, uint32_t *out)
{
	uint32_t words = 0, start = 0;
	while (start < n) {
		uint32_t end = start + 1;
		while (end < n && v[end] == v[start] && end - start < 0xFFFFFFu) end++;
		if (out) out[words] = (end - start) | ((uint32_t)v[start] << 24);
		words++;
		start = end;
	}
	if (out) out[words] = 0;
	return words + 1;
}

/* Decode (rle_decompress :90-116): the first word is expanded unconditionally, then words are consumed
 * until a 0 word (:98-108).  Returns the number of bytes produced (never more than `cap`). */
VO_EXPORT uint32_t vo_rle_decode(const uint32_t *words, uint8_t *out, uint32_t cap)
{
	uint32_t i = 0, b = 0;
	do {
		uint32_t run = words[i] & 0xFFFFFFu;
		if (run > cap - b) run = cap - b;
		memset(out + b, (int)(words[i] >> 24), run);
		b += run;
	} while (words[++i]);
	return b;
}

VO_EXPORT uint64_t vo_fnv1a(const void *p, uint64_t n, uint64_t h)
{
	const uint8_t *b = p;
	if (!h) h = 1469598103934665603ull;
	for (uint64_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}

/* Whole-world rebuild on host threads, same contract as the reference harness' vr_world_rebuild:
 * mode 0 splat / 1 mesh; hashes[k] = FNV-1a 64 of chunk k's output bytes; counts[k*8 + 0..4] = splat
 * items, [5] = vbo items, [6] = ibo items.  Returns wall seconds. */
VO_EXPORT double vo_world_rebuild(const vo_world *w, const uint32_t *ids, uint32_t n_ids, int mode, int nthreads,
                                  uint64_t *hashes, uint32_t *counts)
{
	uint32_t total = 1u << (w->bits[0] + w->bits[1] + w->bits[2]);
	if (!ids) n_ids = total;
	if (nthreads <= 0) nthreads = omp_get_max_threads();
	size_t N = (size_t)1 << (3 * w->rb);
	double t0 = omp_get_wtime();
	#pragma omp parallel num_threads(nthreads)
	{
		int16_t *geom = malloc(N * 100);
		uint32_t *idx = malloc(N * 80);
		#pragma omp for schedule(dynamic, 1)
		for (uint32_t k = 0; k < n_ids; k++) {
			uint32_t id = ids ? ids[k] : k, it[8] = {0};
			uint64_t h = 0;                /* hashes == NULL: a timing run, nothing is hashed */
			if (mode == 0) {
				uint32_t n = vo_chunk_splat(w, id, geom, it);
				if (hashes) h = vo_fnv1a(geom, (uint64_t)n * 2, 0);
			} else {
				vo_chunk_mesh(w, id, geom, idx, &it[5], &it[6]);
				if (hashes) {
					h = vo_fnv1a(geom, (uint64_t)it[5] * 2, 0);
					h = vo_fnv1a(idx, (uint64_t)it[6] * 4, h);
				}
			}
			if (hashes) hashes[k] = h;
			if (counts) memcpy(counts + (size_t)k * 8, it, sizeof it);
		}
		free(geom); free(idx);
	}
	return omp_get_wtime() - t0;
}

/*
 * LOD-node aggregation (the gather loops of gfx_update_svl, gfx/vsplat.c:209-323): node `node` of
 * level `lod` (index = flatten3(chunk offset >> lod, max_bitw - min(lod, max_bitw)), :214-229) concatenates the
 * level-`lod` segment of every member chunk's splat list; members are visited x outer, y, z inner (:264-266) and a
 * segment starts after the chunk's lower levels (:297-300).  svl[c] = chunk c's splat list, items = [n_chunks][5].
 * Returns the node's int16 item count (GeometrySVL.vbo_items, :325); `out` may be NULL to only count.
 */
VO_EXPORT uint32_t vo_lod_node(const int32_t bits[3], int32_t lod, uint32_t node, const int16_t *const *svl,
                               const uint32_t *items, int16_t *out)
{
	int32_t ob[3];
	uint32_t o[3], lo[3], hi[3], total = 0;
	for (int i = 0; i < 3; i++) ob[i] = bits[i] - (lod < bits[i] ? lod : bits[i]);
	o[0] = node & ((1u << ob[0]) - 1); o[1] = (node >> ob[0]) & ((1u << ob[1]) - 1); o[2] = node >> (ob[0] + ob[1]);
	for (int i = 0; i < 3; i++) {
		lo[i] = o[i] << lod; hi[i] = (o[i] + 1) << lod;
		if (hi[i] > (1u << bits[i])) hi[i] = 1u << bits[i];
	}
	for (uint32_t x = lo[0]; x < hi[0]; x++) for (uint32_t y = lo[1]; y < hi[1]; y++) for (uint32_t z = lo[2]; z < hi[2]; z++) {
		uint32_t c = ((z << bits[1] | y) << bits[0]) | x, start = 0;
		const uint32_t *it = items + (size_t)c * 5;
		if (!it[lod]) continue;
		for (int l = 0; l < lod; l++) start += it[l];
		if (out) memcpy(out + total, svl[c] + start, (size_t)it[lod] * sizeof(int16_t));
		total += it[lod];
	}
	return total;
}

/* ---- pick ray: chunkset_edit_raycast_until_solid (chunkset/edit.c:248-314) ----
 * The walk restated with every float operation spelled out -- IEEE divide, the sum of squares as one multiply and two
 * fused multiply-adds in the association the reference build uses (fmaf(t2, t2, fmaf(t0, t0, t1 * t1))), float square
 * root, unsigned <-> float conversions that stick at 0xFFFFFFFF for negative / too large values -- so that it takes the
 * same side at every step as the compiled reference; this is the arithmetic the device kernel (vp_edit.cu k_raycast)
 * uses, operation for operation.  `normal` is in/out like the reference's argument.  Returns the voxel hit (0 = none). */
static uint32_t f2u_sat(float f) { return (!(f > -1.0f) || f >= 4294967296.0f) ? 0xFFFFFFFFu : (uint32_t)f; }

#pragma GCC push_options
#pragma GCC optimize ("fp-contract=off")
VO_EXPORT int vo_raycast(const vo_world *w, const float origin[3], const float vector[3], uint32_t coord[3], int8_t normal[3])
{
	float d[3], next[3], step[3];
	for (int i = 0; i < 3; i++) coord[i] = (uint32_t)(int)origin[i];
	for (int i = 0; i < 3; i++) {
		const float t0 = vector[0] / vector[i], t1 = vector[1] / vector[i], t2 = vector[2] / vector[i];
		const float sq = t1 * t1;
		d[i] = sqrtf(fmaf(t2, t2, fmaf(t0, t0, sq)));
		if (0.0f > vector[i]) { step[i] = -1.0f; const float a = origin[i] - (float)coord[i]; next[i] = a * d[i]; }
		else { step[i] = 1.0f; const float a = (float)coord[i] + 1.0f; const float b = a - origin[i]; next[i] = b * d[i]; }
	}
	for (int loops = 4095; loops > 0; loops--) {
		int side = 0;
		if (next[0] > next[1]) side = 1;
		if (next[side] > next[2]) side = 2;
		next[side] = next[side] + d[side];
		coord[side] = f2u_sat((float)coord[side] + step[side]);
		if (coord[0] >= wdim(w, 0) || coord[1] >= wdim(w, 1) || coord[2] >= wdim(w, 2)) continue;      /* edit.c:22-25 */
		const uint8_t v = vox_at(w, coord[0], coord[1], coord[2]);
		if (v) { normal[side] = (-vector[side] > 0.0f) ? 1 : -1; return v; }
	}
	return 0;
}
#pragma GCC pop_options
