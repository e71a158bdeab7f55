// vp_worldfile.cu -- on-disk world snapshot: checkpoint / resume of the resident world.
//
// Layout = the reference's exporter `command_export` (src/deadcode.c:320-350): the magic byte 0x89 and "VOXPLAT",
// root_bitw (1 byte), max_bitw (3 bytes), every chunk's RLE stream in chunk-id order, then the shadow map.  The
// exporter predates the 32-bit RLE words of chunkset/rle.c (it walks the stream two bytes at a time and writes
// shadow_map_length BYTES of the uint16 map), so the stream section here uses what rle_compress produces today --
// `run | value << 24` words with their 0 terminator (rle.c:44-87) -- and the map is written whole,
// (X+Y)*Z uint16 entries (shadow.h:26-51).  Encode and decode run on the device (vp_rle.cu); the host only moves
// the compressed bytes.
#include "vp_internal.h"
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <memory>

namespace {

constexpr unsigned char kMagic[8] = {0x89, 'V', 'O', 'X', 'P', 'L', 'A', 'T'};
constexpr uint32_t kBatch = 4096;                   // chunks per encode / decode call

struct FileCloser { void operator()(FILE *f) const { if (f) fclose(f); } };
using File = std::unique_ptr<FILE, FileCloser>;

bool whole_world(const vp_ctx *c) { return c->cfg.slab_z0 == 0 && c->cfg.slab_z1 == c->nz; }

} // namespace

extern "C" int vp_world_file_info(const char *path, int32_t *root_bitw, int32_t max_bitw[3], uint64_t *file_bytes)
{
	if (!path) return VP_ERR_ARG;
	File f(fopen(path, "rb"));
	if (!f) return VP_ERR_IO;
	unsigned char h[12];
	if (fread(h, 1, 12, f.get()) != 12 || memcmp(h, kMagic, 8) != 0) return VP_ERR_IO;
	if (root_bitw) *root_bitw = h[8];
	if (max_bitw) for (int i = 0; i < 3; i++) max_bitw[i] = h[9 + i];
	if (file_bytes) {
		if (fseek(f.get(), 0, SEEK_END) != 0) return VP_ERR_IO;
		*file_bytes = (uint64_t)ftell(f.get());
	}
	return VP_OK;
}

extern "C" int vp_world_save(vp_ctx *c, const char *path, uint64_t *bytes_written)
{
	if (!c || !path) return vp_fail(c, VP_ERR_ARG, "vp_world_save: null argument");
	if (!whole_world(c)) return vp_fail(c, VP_ERR_ARG, "vp_world_save: the context must hold the whole world (slab = all chunk rows)");
	File f(fopen(path, "wb"));
	if (!f) return vp_fail(c, VP_ERR_IO, "vp_world_save: cannot open the file for writing");
	unsigned char h[12];
	memcpy(h, kMagic, 8);
	h[8] = (unsigned char)c->rb;
	for (int i = 0; i < 3; i++) h[9 + i] = (unsigned char)c->cfg.max_bitw[i];
	if (fwrite(h, 1, 12, f.get()) != 12) return vp_fail(c, VP_ERR_IO, "vp_world_save: write failed");
	uint64_t total = 12;

	const uint32_t n_chunks = (uint32_t)c->nx * c->ny * c->nz;
	std::vector<uint32_t> ids, words;
	std::vector<uint64_t> offs;
	uint32_t batch = kBatch;
	for (uint32_t first = 0; first < n_chunks;) {
		const uint32_t n = std::min(batch, n_chunks - first);
		ids.resize(n); offs.assign((size_t)n + 1, 0);
		for (uint32_t i = 0; i < n; i++) ids[i] = first + i;
		// the first call sizes the streams (word_offsets is filled even when the host buffer is too small), the second
		// fetches them; word_offsets[n] == 0 after a failure means the device arena overflowed: retry with half the batch
		int rc = vp_encode_chunks_rle(c, ids.data(), n, words.data(), words.size(), offs.data());
		if (rc == VP_ERR_ARENA_FULL && offs[n] > words.size()) {
			words.resize(offs[n]);
			rc = vp_encode_chunks_rle(c, ids.data(), n, words.data(), words.size(), offs.data());
		}
		if (rc == VP_ERR_ARENA_FULL && offs[n] == 0 && n > 1) { batch = n / 2; continue; }
		if (rc) return rc;
		if (fwrite(words.data(), 4, offs[n], f.get()) != offs[n]) return vp_fail(c, VP_ERR_IO, "vp_world_save: write failed");
		total += offs[n] * 4;
		first += n;
	}

	const uint32_t shw = (uint32_t)((c->nx + c->ny) << c->rb), Z = (uint32_t)c->nz << c->rb;
	std::vector<uint16_t> rows((size_t)shw * std::min<uint32_t>(Z, 256));
	for (uint32_t z = 0; z < Z;) {
		const uint32_t zn = std::min<uint32_t>(256, Z - z);
		int rc = vp_download_shadow_rows(c, z, z + zn, rows.data());
		if (rc) return rc;
		if (fwrite(rows.data(), 2, (size_t)zn * shw, f.get()) != (size_t)zn * shw) return vp_fail(c, VP_ERR_IO, "vp_world_save: write failed");
		total += (uint64_t)zn * shw * 2;
		z += zn;
	}
	if (fflush(f.get()) != 0) return vp_fail(c, VP_ERR_IO, "vp_world_save: write failed");
	if (bytes_written) *bytes_written = total;
	return VP_OK;
}

extern "C" int vp_world_load(vp_ctx *c, const char *path)
{
	if (!c || !path) return vp_fail(c, VP_ERR_ARG, "vp_world_load: null argument");
	if (!whole_world(c)) return vp_fail(c, VP_ERR_ARG, "vp_world_load: the context must hold the whole world (slab = all chunk rows)");
	File f(fopen(path, "rb"));
	if (!f) return vp_fail(c, VP_ERR_IO, "vp_world_load: cannot open the file");
	unsigned char h[12];
	if (fread(h, 1, 12, f.get()) != 12 || memcmp(h, kMagic, 8) != 0) return vp_fail(c, VP_ERR_IO, "vp_world_load: not a VOXPLAT world file");
	if (h[8] != c->rb || h[9] != c->cfg.max_bitw[0] || h[10] != c->cfg.max_bitw[1] || h[11] != c->cfg.max_bitw[2])
		return vp_fail(c, VP_ERR_ARG, "vp_world_load: the file's root_bitw / max_bitw differ from the context's");

	// The streams carry no length prefix: read the rest of the file and cut it at the 0 terminators.
	if (fseek(f.get(), 0, SEEK_END) != 0) return vp_fail(c, VP_ERR_IO, "vp_world_load: seek failed");
	const uint64_t fsize = (uint64_t)ftell(f.get());
	if (fseek(f.get(), 12, SEEK_SET) != 0 || fsize < 12 || (fsize - 12) % 4) return vp_fail(c, VP_ERR_IO, "vp_world_load: file size is not a whole number of words");
	std::vector<uint32_t> buf((fsize - 12) / 4);
	if (fread(buf.data(), 4, buf.size(), f.get()) != buf.size()) return vp_fail(c, VP_ERR_IO, "vp_world_load: read failed");
	const uint32_t n_chunks = (uint32_t)c->nx * c->ny * c->nz;
	const uint32_t shw = (uint32_t)((c->nx + c->ny) << c->rb), Z = (uint32_t)c->nz << c->rb;
	const size_t map_words = (size_t)shw * Z / 2;
	if (buf.size() < map_words + 2 * (size_t)n_chunks) return vp_fail(c, VP_ERR_IO, "vp_world_load: the file is truncated");
	const size_t stream_words = buf.size() - map_words;
	std::vector<uint32_t> ids;
	std::vector<uint64_t> offs;
	size_t pos = 0;
	uint32_t batch = kBatch;
	for (uint32_t first = 0; first < n_chunks;) {
		const uint32_t n = std::min(batch, n_chunks - first);
		const size_t start = pos;
		ids.resize(n); offs.resize((size_t)n + 1);
		for (uint32_t i = 0; i < n; i++) {
			ids[i] = first + i;
			offs[i] = pos;
			while (pos < stream_words && buf[pos] != 0) pos++;
			if (pos >= stream_words) return vp_fail(c, VP_ERR_IO, "vp_world_load: the file ends inside a chunk stream");
			pos++;
		}
		offs[n] = pos;
		int rc = vp_upload_chunks_rle(c, ids.data(), n, buf.data(), offs.data());
		if (rc == VP_ERR_ARENA_FULL && n > 1) { batch = n / 2; pos = start; continue; }     // device staging too small: smaller batch
		if (rc) return rc;
		first += n;
	}
	if (pos != stream_words) return vp_fail(c, VP_ERR_IO, "vp_world_load: stream section and shadow map do not add up to the file size");
	return vp_upload_shadow_rows(c, 0, Z, reinterpret_cast<const uint16_t *>(buf.data() + pos));
}
