// vp_rle.cu -- device RLE codec, byte-identical to the reference's rle.c.
//   word = run (24 bit) | value << 24, stream terminated by a 0 word (rle.c:17-26)
//   decode = rle_decompress (rle.c:90-116), encode = rle_compress (rle.c:44-87)
// Both kernels run one CTA per stream and walk it in tiles; ordering inside a stream comes from block
// prefix scans (warp shuffles), never from atomics.
#include "vp_device.cuh"
#include <cstring>
#include <algorithm>
using namespace vp;

namespace {

constexpr int kT = 256;                  // threads per CTA
constexpr int kDecTile = 2048;           // run words per decode tile
constexpr int kEncTile = kT * 16;        // bytes per encode tile

// exclusive block scan of one value per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t *wsum, uint32_t *total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
	__syncthreads();                      // protect wsum from the previous use
	if (lane == 31) wsum[warp] = inc;
	__syncthreads();
	uint32_t pre = inc - v, tot = 0;
	#pragma unroll
	for (int k = 0; k < kT / 32; k++) { uint32_t s = wsum[k]; if (k < warp) pre += s; tot += s; }
	*total = tot;
	return pre;
}

// ---- decode ----------------------------------------------------------------------------------------
// Output-parallel expansion: the run lengths of a tile of words are prefix-summed into start offsets;
// every thread then produces whole 16-byte groups of the output (one uint4 store each), locating the
// run that covers the group's first byte by binary search in shared memory and walking forward.
// A group straddling two word tiles is produced by the later tile, which starts at the run that
// contains the first unwritten byte.
__global__ void __launch_bounds__(kT)
k_rle_decode(const uint32_t *__restrict__ words, const unsigned long long *__restrict__ offsets,
             const int32_t *__restrict__ slots, uint8_t *__restrict__ dst_base, uint32_t N, uint32_t *__restrict__ status)
{
	__shared__ uint32_t s_start[kDecTile + 1];
	__shared__ uint8_t s_val[kDecTile];
	__shared__ uint32_t s_wsum[kT / 32];
	__shared__ uint32_t s_next[2];
	const int i = blockIdx.x, tid = threadIdx.x;
	if (slots && slots[i] < 0) return;
	uint8_t *dst = dst_base + (size_t)(slots ? slots[i] : i) * N;
	const uint32_t *w = words + offsets[i];
	const uint32_t nw = (uint32_t)(offsets[i + 1] - offsets[i]) - 1u;       // runs, without the terminator
	uint32_t widx = 0, wstart = 0, pos = 0, bad = 0;
	constexpr int IPT = kDecTile / kT;
	while (widx < nw) {
		const uint32_t tc = min((uint32_t)kDecTile, nw - widx);
		uint32_t c[IPT], sum = 0;
		#pragma unroll
		for (int k = 0; k < IPT; k++) {
			const uint32_t j = tid * IPT + k;
			uint32_t word = j < tc ? __ldg(w + widx + j) : 0u;
			c[k] = word & 0xFFFFFFu;
			if (c[k] > N) { c[k] = N; bad = 1u; }          // no run of a valid stream is longer than the chunk
			if (j < tc) s_val[j] = (uint8_t)(word >> 24);
			sum += c[k];
		}
		// a thread's 8 runs are worth at most N + 1 to the block scan: the sum of a (malformed) tile then stays far below
		// 2^32 (256 (N + 1)), and every start offset past the saturation point is > N, i.e. never searched for
		uint32_t total, pre = block_scan_excl(min(sum, N + 1u), s_wsum, &total) + wstart;
		#pragma unroll
		for (int k = 0; k < IPT; k++) { const uint32_t j = tid * IPT + k; if (j <= tc) s_start[j] = pre; pre += c[k]; }
		if (tid == kT - 1) s_start[tc] = wstart + total;      // (also covers tc == kDecTile)
		__syncthreads();
		const bool last = widx + tc == nw;
		uint32_t covered = min(s_start[tc], N);
		if (last && s_start[tc] != N && tid == 0) atomicExch(status, 1u);         // stream length != chunk volume
		const uint32_t g_end = covered / 16;
		for (uint32_t g = pos / 16 + tid; g < g_end; g += kT) {
			const uint32_t o = g * 16;
			uint32_t lo = 0, hi = tc;                  // largest k with s_start[k] <= o
			while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (s_start[mid] <= o) lo = mid; else hi = mid; }
			uint32_t k = lo;
			if (s_start[k + 1] >= o + 16) {            // one run covers the whole group (air, solid interior): broadcast
				const uint32_t v = (uint32_t)s_val[k] * 0x01010101u;
				*reinterpret_cast<uint4 *>(dst + o) = make_uint4(v, v, v, v);
				continue;
			}
			uint32_t out[4] = {0, 0, 0, 0};
			#pragma unroll
			for (int b = 0; b < 16; b++) {
				while (k + 1 < tc && o + b >= s_start[k + 1]) k++;
				out[b >> 2] |= (uint32_t)s_val[k] << ((b & 3) * 8);
			}
			*reinterpret_cast<uint4 *>(dst + o) = make_uint4(out[0], out[1], out[2], out[3]);
		}
		// next tile starts at the run containing the first unwritten byte
		if (tid == 0) {
			const uint32_t np = g_end * 16;
			if (np >= s_start[tc]) { s_next[0] = tc; s_next[1] = s_start[tc]; }
			else {
				uint32_t lo = 0, hi = tc;
				while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (s_start[mid] <= np) lo = mid; else hi = mid; }
				if (lo == 0 && tc == kDecTile) { lo = tc; }     // cannot happen for valid data (a run cannot exceed a chunk); avoid livelock
				s_next[0] = lo; s_next[1] = s_start[lo];
			}
		}
		__syncthreads();
		pos = g_end * 16;
		const uint32_t adv = s_next[0];
		wstart = s_next[1];
		widx += adv;
		if (covered >= N && !last && tid == 0) atomicExch(status, 1u);       // runs left after the chunk is full: stream too long
		if (covered >= N || last) break;           // a short final tile: the tail below zero-fills and flags the stream
		if (adv == 0) { if (tid == 0) atomicExch(status, 1u); break; }      // no progress (zero-length runs): malformed
		__syncthreads();
	}
	// a short stream leaves the tail undefined in the reference; make it deterministic (zeros) and flag it
	for (uint32_t o = pos + tid * 16; o < N; o += kT * 16) *reinterpret_cast<uint4 *>(dst + o) = make_uint4(0, 0, 0, 0);
	if (pos < N && tid == 0) atomicExch(status, 1u);
	if (bad) atomicExch(status, 1u);
}

// ---- encode ----------------------------------------------------------------------------------------
// Head flags (v[i] != v[i-1]) per 16-byte group -> popc -> block scan gives every run its word index
// (stable order); run length = distance to the next head, resolved through a shared-memory list of head
// positions; the last head of a tile is carried into the next tile.  Two passes over the chunk (count,
// then emit into a region reserved with one atomicAdd); the second pass hits L2.
__device__ __forceinline__ uint32_t head_mask16(uint4 v, uint32_t prev_byte, bool first)
{
	// shifted-by-one-byte copy of the 16 bytes, then per-byte inequality
	uint32_t s0 = (v.x << 8) | prev_byte, s1 = (v.y << 8) | (v.x >> 24), s2 = (v.z << 8) | (v.y >> 24), s3 = (v.w << 8) | (v.z >> 24);
	uint32_t m = nz4(v.x ^ s0) | (nz4(v.y ^ s1) << 4) | (nz4(v.z ^ s2) << 8) | (nz4(v.w ^ s3) << 12);
	return first ? (m | 1u) : m;
}

__global__ void __launch_bounds__(kT)
k_rle_encode(const uint8_t *__restrict__ src_base, const int32_t *__restrict__ slots, uint32_t N,
             uint32_t *__restrict__ arena_words, VpArenaDev *__restrict__ st,
             unsigned long long *__restrict__ out_offsets, uint32_t *__restrict__ out_counts)
{
	__shared__ uint32_t s_hp[kEncTile];
	__shared__ uint8_t s_hv[kEncTile];
	__shared__ uint32_t s_wsum[kT / 32];
	__shared__ unsigned long long s_off;
	const int i = blockIdx.x, tid = threadIdx.x;
	if (slots && slots[i] < 0) { if (tid == 0) { out_counts[i] = 0; out_offsets[i] = 0; } return; }
	const uint8_t *src = src_base + (size_t)(slots ? slots[i] : i) * N;

	// pass 1: count heads
	uint32_t cnt = 0;
	for (uint32_t o = tid * 16; o < N; o += kT * 16) {
		uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + o));
		uint32_t prev = o ? (uint32_t)__ldg(src + o - 1) : 0u;
		cnt += __popc(head_mask16(v, prev, o == 0));
	}
	uint32_t runs;
	block_scan_excl(cnt, s_wsum, &runs);
	if (tid == 0) {
		const unsigned long long bytes = ((unsigned long long)runs + 1ull) * 4ull;
		unsigned long long off = atomicAdd(&st->cursor, bytes);
		if (off + bytes > st->capacity) { atomicExch(&st->overflow, 1u); off = ~0ull; }
		s_off = off;
		out_counts[i] = runs + 1u;
		out_offsets[i] = off == ~0ull ? ~0ull : off / 4ull;
	}
	__syncthreads();
	if (s_off == ~0ull) return;
	uint32_t *out = arena_words + s_off / 4ull;

	// pass 2: emit
	uint32_t gbase = 0, carry_pos = 0, carry_val = 0;
	for (uint32_t t0 = 0; t0 < N; t0 += kEncTile) {
		const uint32_t o = t0 + tid * 16;
		uint32_t m = 0; uint4 v = make_uint4(0, 0, 0, 0);
		if (o < N) {
			v = __ldg(reinterpret_cast<const uint4 *>(src + o));
			uint32_t prev = o ? (uint32_t)__ldg(src + o - 1) : 0u;
			m = head_mask16(v, prev, o == 0);
		}
		uint32_t total, pre = block_scan_excl(__popc(m), s_wsum, &total);
		const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
		while (m) {
			const int b = __ffs(m) - 1; m &= m - 1;
			s_hp[pre] = o + b;
			s_hv[pre] = (uint8_t)(vv[b >> 2] >> ((b & 3) * 8));
			pre++;
		}
		__syncthreads();
		if (total) {
			if (gbase && tid == 0) out[gbase - 1] = (s_hp[0] - carry_pos) | (carry_val << 24);
			for (uint32_t k = tid; k + 1 < total; k += kT) out[gbase + k] = (s_hp[k + 1] - s_hp[k]) | ((uint32_t)s_hv[k] << 24);
			carry_pos = s_hp[total - 1]; carry_val = s_hv[total - 1];
			gbase += total;
		}
		__syncthreads();
	}
	if (tid == 0) { out[gbase - 1] = (N - carry_pos) | (carry_val << 24); out[gbase] = 0u; }
}

} // namespace

cudaError_t vp_launch_rle_decode(const uint32_t *d_words, const unsigned long long *d_offsets, const int32_t *d_slots,
                                 uint32_t n, uint8_t *dst_base, uint32_t N, uint32_t *d_status, cudaStream_t s)
{
	if (!n) return cudaSuccess;
	k_rle_decode<<<n, kT, 0, s>>>(d_words, d_offsets, d_slots, dst_base, N, d_status);
	return cudaGetLastError();
}

cudaError_t vp_launch_rle_encode(const uint8_t *src_base, const int32_t *d_slots, uint32_t n, uint32_t N, uint32_t *d_arena_words,
                                 VpArenaDev *state, unsigned long long *d_offsets, uint32_t *d_counts, cudaStream_t s)
{
	if (!n) return cudaSuccess;
	k_rle_encode<<<n, kT, 0, s>>>(src_base, d_slots, N, d_arena_words, state, d_offsets, d_counts);
	return cudaGetLastError();
}
