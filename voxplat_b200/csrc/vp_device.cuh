// vp_device.cuh -- device-side helpers shared by the kernels: PTX wrappers for mbarrier + TMA bulk
// copies (cp.async.bulk -> SASS UBLKCP), byte->bit packing, bit-pair compression, shadow sampling.
#pragma once
#include "vp_internal.h"

namespace vp {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok, a = smem_u32(bar);
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok) : "r"(a), "r"(parity), "r"(2000u) : "memory");      // suspend up to ~2 us instead of spinning
	} while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16 B aligned).
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 4 bytes -> 4 bits (bit k = byte k != 0).
__device__ __forceinline__ uint32_t nz4(uint32_t w)
{
	uint32_t t = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;    // bit 7 of each byte = byte != 0
	return (t * 0x00204081u) >> 28;                                        // gather bits 7,15,23,31 -> 0..3 (top nibble of the product)
}
// 16 bytes -> 16 bits.
__device__ __forceinline__ uint32_t nz16(uint4 v)
{
	return nz4(v.x) | (nz4(v.y) << 4) | (nz4(v.z) << 8) | (nz4(v.w) << 12);
}

// Bit 7 of each byte = byte != 0 (other bits cleared).
__device__ __forceinline__ uint32_t nzflags(uint32_t w)
{
	return (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
}
// 16 bytes -> 128 * (16-bit mask) split over two dot-product accumulators: the byte flags (0 / 0x80) are weighted
// 1,2,4,..,128 by IDP.4A, so the gather runs on the dot-product unit instead of multiply + shift + or chains.
// Returns 128 * mask(bytes 0..7) + 256 * 128 * mask(bytes 8..15)  (< 2^23).
__device__ __forceinline__ uint32_t nz16x128(uint4 v)
{
	uint32_t a = __dp4a(nzflags(v.x), 0x08040201u, 0u);
	a = __dp4a(nzflags(v.y), 0x80402010u, a);
	uint32_t b = __dp4a(nzflags(v.z), 0x08040201u, 0u);
	b = __dp4a(nzflags(v.w), 0x80402010u, b);
	return a + (b << 8);
}
// Bulk copy shared -> global (bytes % 16 == 0, both addresses 16 B aligned), tracked by the thread's bulk async-group.
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the committed bulk copies have READ their shared-memory source (it may then be reused / the CTA may exit)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy writes to shared memory visible to the async proxy (bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// L2 prefetch of a byte range (no shared-memory destination, no completion tracking).
__device__ __forceinline__ void l2_prefetch(const void *gsrc, uint32_t bytes)
{
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// OR adjacent bit pairs and pack the 32 results into the low half: out bit k = in bit 2k | in bit 2k+1.
__device__ __forceinline__ uint64_t pair_or_compress(uint64_t a)
{
	uint64_t b = (a | (a >> 1)) & 0x5555555555555555ull;
	b = (b | (b >> 1)) & 0x3333333333333333ull;
	b = (b | (b >> 2)) & 0x0f0f0f0f0f0f0f0full;
	b = (b | (b >> 4)) & 0x00ff00ff00ff00ffull;
	b = (b | (b >> 8)) & 0x0000ffff0000ffffull;
	b = (b | (b >> 16)) & 0x00000000ffffffffull;
	return b;
}

// shadow_sample / shadow_sample_normal (reference shadow.h:45-75) on the device copy of the map.
// second = +1 (sample) or -1 (sample_normal).  All arithmetic is uint32 like the reference; y+1 == 0
// can only compare false, so the result is 1 without a load (frozen reference behaviours).
__device__ __forceinline__ int shadow_pair(const VpWorldDev &w, uint32_t x, uint32_t y, uint32_t z, int second)
{
	uint32_t lim = y + 1u;
	if (lim == 0) return 1;
	size_t idx = (size_t)(x + y) + (size_t)w.sh_w * (size_t)(z - w.sh_z0);
	uint32_t a = __ldg(w.shadow + idx), b = __ldg(w.shadow + idx + second);
	return !(a < lim && b < lim);
}

// Slot of chunk (cx,cy,cz); -1 for null chunks, chunks outside the world (mesher.c:393-394) and rows not
// held by this device.
__device__ __forceinline__ int chunk_slot(const VpWorldDev &w, int cx, int cy, int cz)
{
	if ((unsigned)cx >= (1u << w.bits[0]) || (unsigned)cy >= (1u << w.bits[1]) || (unsigned)cz >= (1u << w.bits[2])) return -1;
	if (cz < w.ez0 || cz >= w.ez1) return -1;
	return __ldg(w.slot + ((((size_t)(cz - w.ez0) << w.bits[1]) | (unsigned)cy) << w.bits[0] | (unsigned)cx));
}

} // namespace vp
