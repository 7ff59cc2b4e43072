// engine.cu — libb200compact.so: context lifecycle, CRC tables, device scan, and the chunk-codec entry points of
// include/b200c.h (b200c_compress_chunks / b200c_decompress_chunks / ICompressor single-buffer calls).
// No CPU fallback: every compute entry point needs a CUDA device and fails with B200C_ECUDA otherwise.
#include "engine.cuh"
#include "scan.cuh"
#include "codec.cuh"
#include <climits>
#include <vector>

using namespace b200c;

namespace b200c {

enum { WSC_IN = 0, WSC_OUT, WSC_SLOTS, WSC_FILELEN, WSC_SEGRAW, WSC_OFFS, WSC_ACC, WSC_ERR, WSC_CHOFFS, WSC_SCAN0, WSC_SCAN1, WSC_SCAN2 };

static void build_tables(DevTables* t) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        t->crc_t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int k = 1; k < 4; k++) t->crc_t[k][i] = (t->crc_t[k - 1][i] >> 8) ^ t->crc_t[0][t->crc_t[k - 1][i] & 0xff];
    // x^(8*2^k)
    uint32_t sq = 0x00800000u;          // x^8, reflected
    for (int k = 0; k < 64; k++) { t->xp_pow2[k] = sq; sq = gf2_mulmod(sq, sq); }
    auto xpow8n = [&](uint64_t n) { uint32_t r = 0x80000000u; for (int k = 0; n; k++, n >>= 1) if (n & 1) r = gf2_mulmod(r, t->xp_pow2[k]); return r; };
    uint32_t x128 = xpow8n(128);
    for (int j = 0; j < 4; j++)
        for (uint32_t b = 0; b < 256; b++) t->crc_adv128[j][b] = gf2_mulmod(b << (8 * j), x128);
    for (int l = 0; l < 32; l++) t->xp_lane[l] = xpow8n(4 * (32 - l));
}

template <typename TIn>
int exclusive_scan(b200c_ctx* c, const TIn* in, uint64_t n, uint64_t* out, int slot, int depth) {
    if (n == 0) { B200C_CUDA_TRY(c, cudaMemsetAsync(out, 0, sizeof(uint64_t), c->stream)); return B200C_OK; }
    uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) { B200C_LAUNCH(c, k_scan_down<TIn>, 1, SCAN_THREADS, 0, in, n, (const uint64_t*)nullptr, out, out + n); return B200C_OK; }
    if (depth > 2) { c->err = "scan recursion too deep"; return B200C_EINVAL; }
    uint64_t* sums; B200C_TRY(ws_typed(c, slot + depth, tiles + 1, &sums));
    B200C_LAUNCH(c, k_scan_reduce<TIn>, (unsigned)tiles, SCAN_THREADS, 0, in, n, sums);
    B200C_TRY(exclusive_scan<uint64_t>(c, sums, tiles, sums, slot, depth + 1));
    B200C_LAUNCH(c, k_scan_down<TIn>, (unsigned)tiles, SCAN_THREADS, 0, in, n, (const uint64_t*)sums, out, out + n);
    return B200C_OK;
}
template int exclusive_scan<uint32_t>(b200c_ctx*, const uint32_t*, uint64_t, uint64_t*, int, int);
template int exclusive_scan<uint64_t>(b200c_ctx*, const uint64_t*, uint64_t, uint64_t*, int, int);

// device-side chunk compression of a resident stream, in two halves so that the LCS writer can compress a window, look at the
// chunk sizes, and only then decide where the file ends:
//   compress_slots_device: every chunk of d_in[0..n) -> its fixed-stride slot (+ CRC), file_len[i] = bytes + 4, seg_raw[i] for the digest
//   pack_digest_device:    first nchunks slots -> dense Data.db image, chunk offsets (nchunks + 1), Digest.crc32
int compress_slots_device(b200c_ctx* c, int comp, const uint8_t* d_in, uint64_t n, int chunk_len, int max_clen,
                          uint8_t* slots, int stride, uint32_t* file_len, uint32_t* seg_raw) {
    uint64_t nchunks = (n + chunk_len - 1) / chunk_len;
    if (!nchunks) return B200C_OK;
    int tab_bytes = comp == COMP_SNAPPY ? 32768 : 16384;
    size_t smem = (size_t)tab_bytes + chunk_len + 16;
    B200C_LAUNCH(c, k_compress_chunks, (unsigned)nchunks, 32, smem, c->d_tables, comp, tab_bytes, d_in, n, chunk_len, max_clen, slots, stride, file_len, seg_raw);
    return B200C_OK;
}
int pack_digest_device(b200c_ctx* c, const uint8_t* slots, int stride, const uint32_t* file_len, const uint32_t* seg_raw, uint64_t nchunks,
                       uint8_t* d_out, uint64_t out_cap, uint64_t* d_offs, uint64_t* out_len, uint32_t* digest, int ws_base) {
    if (!nchunks) { *out_len = 0; *digest = 0; B200C_CUDA_TRY(c, cudaMemsetAsync(d_offs, 0, 8, c->stream)); return B200C_OK; }
    uint32_t* acc; B200C_TRY(ws_typed(c, ws_base + WSC_ACC, 4, &acc));
    B200C_CUDA_TRY(c, cudaMemsetAsync(acc, 0, 16, c->stream));
    B200C_TRY(exclusive_scan<uint32_t>(c, file_len, nchunks, d_offs, ws_base + WSC_SCAN0, 0));
    uint64_t* h = (uint64_t*)c->h_pinned;                 // the total size must be known on the host before packing into the caller's buffer
    B200C_CUDA_TRY(c, cudaMemcpyAsync(h, d_offs + nchunks, 8, cudaMemcpyDeviceToHost, c->stream));
    B200C_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *out_len = h[0];
    if (h[0] > out_cap) { c->err = "output buffer too small"; return B200C_ETOOSMALL; }
    B200C_LAUNCH(c, k_pack_chunks, (unsigned)((nchunks + 3) / 4), 128, 0, slots, stride, file_len, d_offs, nchunks, d_out);
    B200C_LAUNCH(c, k_digest, (unsigned)((nchunks + 255) / 256), 256, 0, c->d_tables, seg_raw, d_offs, nchunks, acc);
    B200C_LAUNCH(c, k_digest_final, 1, 1, 0, c->d_tables, d_offs, nchunks, acc);
    uint32_t* h32 = (uint32_t*)c->h_pinned;
    B200C_CUDA_TRY(c, cudaMemcpyAsync(h32, acc + 1, 4, cudaMemcpyDeviceToHost, c->stream));
    B200C_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *digest = h32[0];
    return B200C_OK;
}
int compress_stream_device(b200c_ctx* c, int comp, const uint8_t* d_in, uint64_t n, int chunk_len, int max_clen,
                           uint8_t* d_out, uint64_t out_cap, uint64_t* d_offs /*nchunks+1*/, uint64_t* out_len, uint32_t* digest, int ws_base) {
    uint64_t nchunks = (n + chunk_len - 1) / chunk_len;
    if (nchunks > 0x7fffffffull) { c->err = "too many chunks"; return B200C_EINVAL; }
    int stride = chunk_slot_stride(comp, chunk_len);
    uint8_t* slots = nullptr; uint32_t* file_len = nullptr; uint32_t* seg_raw = nullptr;
    B200C_TRY(ws_typed(c, ws_base + WSC_SLOTS, nchunks * (uint64_t)stride, &slots));
    B200C_TRY(ws_typed(c, ws_base + WSC_FILELEN, nchunks + 1, &file_len));
    B200C_TRY(ws_typed(c, ws_base + WSC_SEGRAW, nchunks + 1, &seg_raw));
    B200C_TRY(compress_slots_device(c, comp, d_in, n, chunk_len, max_clen, slots, stride, file_len, seg_raw));
    return pack_digest_device(c, slots, stride, file_len, seg_raw, nchunks, d_out, out_cap, d_offs, out_len, digest, ws_base);
}

// slice-relative chunk offsets -> absolute offsets in the file image; bases[0] = offset of this slice, bases[1] := offset of the next
__global__ void __launch_bounds__(256) k_offs_add_base(const uint64_t* __restrict__ rel, uint64_t count, uint64_t* __restrict__ bases, uint64_t* __restrict__ offs_abs) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t b = bases[0];
    if (i <= count) offs_abs[i] = b + rel[i];
    if (i == count) bases[1] = b + rel[count];
}

// Same result as compress_stream_device followed by a device->host copy of the file image, but in slices: slice s is copied to the
// caller's host buffer on the copy stream while slice s+1 is being compressed, so that the PCIe read-back hides behind K5.
// d_img: device staging of the whole image (cap img_cap); h_out/h_cap: the caller's buffer. When the image does not fit h_cap the
// copies stop, *out_len is still the full size and the caller reports B200C_ETOOSMALL.
int compress_stream_to_host(b200c_ctx* c, int comp, const uint8_t* d_in, uint64_t n, int chunk_len, int max_clen,
                            uint8_t* d_img, uint64_t img_cap, uint8_t* h_out, uint64_t h_cap, uint64_t* d_offs /*nchunks+1*/,
                            uint64_t* out_len, uint32_t* digest, int ws_base) {
    const uint64_t nchunks = (n + chunk_len - 1) / chunk_len;
    if (nchunks > 0x7fffffffull) { c->err = "too many chunks"; return B200C_EINVAL; }
    if (!nchunks) { *out_len = 0; *digest = 0; B200C_CUDA_TRY(c, cudaMemsetAsync(d_offs, 0, 8, c->stream)); return B200C_OK; }
    enum { MAX_SLICES = 16, MIN_SLICE_CHUNKS = 8192 };
    const int nslices = (int)std::min<uint64_t>(MAX_SLICES, std::max<uint64_t>(1, nchunks / MIN_SLICE_CHUNKS));
    const uint64_t per = (nchunks + nslices - 1) / nslices;
    const int stride = chunk_slot_stride(comp, chunk_len);
    uint8_t* slots; uint32_t *file_len, *seg_raw, *acc; uint64_t *rel, *bases;
    B200C_TRY(ws_typed(c, ws_base + WSC_SLOTS, nchunks * (uint64_t)stride, &slots));
    B200C_TRY(ws_typed(c, ws_base + WSC_FILELEN, nchunks + 1, &file_len));
    B200C_TRY(ws_typed(c, ws_base + WSC_SEGRAW, nchunks + 1, &seg_raw));
    B200C_TRY(ws_typed(c, ws_base + WSC_IN, per + 2, &rel));
    B200C_TRY(ws_typed(c, ws_base + WSC_CHOFFS, MAX_SLICES + 2, &bases));
    B200C_TRY(ws_typed(c, ws_base + WSC_ACC, 4, &acc));
    {   // scan scratch is sized by the first slice; allocate it before the loop so that no slot grows (and synchronises) mid-pipeline
        uint64_t* tmp; uint64_t tiles = (per + SCAN_TILE - 1) / SCAN_TILE;
        B200C_TRY(ws_typed(c, ws_base + WSC_SCAN0, tiles + 1, &tmp));
        B200C_TRY(ws_typed(c, ws_base + WSC_SCAN0 + 1, (tiles + SCAN_TILE - 1) / SCAN_TILE + 1, &tmp));
    }
    B200C_CUDA_TRY(c, cudaMemsetAsync(acc, 0, 16, c->stream));
    B200C_CUDA_TRY(c, cudaMemsetAsync(bases, 0, 8, c->stream));
    uint64_t* h = (uint64_t*)c->h_pinned + 1024;          // slice end offsets land here
    uint64_t copied = 0; bool fits = true;
    auto drain = [&](int s) -> int {                      // slice s is packed once ev_in[1 + s] fires: hand its bytes to the copy engine
        B200C_CUDA_TRY(c, cudaEventSynchronize(c->ev_in[1 + s]));
        uint64_t end = h[s];
        if (end > img_cap) { c->err = "internal error: compressed image exceeds its bound"; return B200C_ECUDA; }
        if (end > h_cap) fits = false;
        if (fits && end > copied) {
            B200C_CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_in[1 + s], 0));
            B200C_CUDA_TRY(c, cudaMemcpyAsync(h_out + copied, d_img + copied, end - copied, cudaMemcpyDeviceToHost, c->copy_stream));
        }
        copied = end;
        return B200C_OK;
    };
    for (int s = 0; s < nslices; s++) {
        const uint64_t a = (uint64_t)s * per, b = std::min<uint64_t>(nchunks, a + per);
        if (a >= b) { h[s] = s ? h[s - 1] : 0; cudaEventRecord(c->ev_in[1 + s], c->stream); continue; }
        const uint64_t nb = std::min<uint64_t>(n, b * (uint64_t)chunk_len) - a * (uint64_t)chunk_len;
        B200C_TRY(compress_slots_device(c, comp, d_in + a * (uint64_t)chunk_len, nb, chunk_len, max_clen, slots + a * (uint64_t)stride, stride, file_len + a, seg_raw + a));
        B200C_TRY(exclusive_scan<uint32_t>(c, file_len + a, b - a, rel, ws_base + WSC_SCAN0, 0));
        B200C_LAUNCH(c, k_offs_add_base, (unsigned)((b - a + 1 + 255) / 256), 256, 0, rel, b - a, bases + s, d_offs + a);
        B200C_LAUNCH(c, k_pack_chunks, (unsigned)((b - a + 3) / 4), 128, 0, slots + a * (uint64_t)stride, stride, file_len + a, d_offs + a, b - a, d_img);
        B200C_CUDA_TRY(c, cudaMemcpyAsync(h + s, bases + s + 1, 8, cudaMemcpyDeviceToHost, c->stream));
        B200C_CUDA_TRY(c, cudaEventRecord(c->ev_in[1 + s], c->stream));
        if (s) B200C_TRY(drain(s - 1));                   // after slice s is queued, so the GPU never waits for the host
    }
    B200C_LAUNCH(c, k_digest, (unsigned)((nchunks + 255) / 256), 256, 0, c->d_tables, seg_raw, d_offs, nchunks, acc);
    B200C_LAUNCH(c, k_digest_final, 1, 1, 0, c->d_tables, d_offs, nchunks, acc);
    B200C_CUDA_TRY(c, cudaMemcpyAsync(h + MAX_SLICES, acc + 1, 4, cudaMemcpyDeviceToHost, c->stream));
    B200C_TRY(drain(nslices - 1));
    B200C_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *out_len = copied; *digest = (uint32_t)h[MAX_SLICES];
    return B200C_OK;
}

int decompress_stream_device(b200c_ctx* c, int comp, const uint8_t* d_data, uint64_t data_len, const uint64_t* d_offs, uint64_t nchunks,
                             int chunk_len, int max_clen, uint64_t data_length, uint8_t* d_out, int verify, ChunkErr* d_err) {
    if (nchunks == 0) return B200C_OK;
    B200C_LAUNCH(c, k_decompress_chunks, (unsigned)((nchunks + 1) / 2), 64, 0, c->d_tables, comp, d_data, data_len, d_offs, nchunks,
                 chunk_len, max_clen, data_length, d_out, verify, d_err);
    return B200C_OK;
}

} // namespace b200c

// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

int b200c_abi_version(void) { return B200C_ABI_VERSION; }

int b200c_device_count(void) {
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n;
}

b200c_ctx* b200c_create(int device, size_t workspace_bytes) {
    int n = b200c_device_count();
    if (device < 0 || device >= n) return nullptr;       // no device => no context: the product path has no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    b200c_ctx* c = new b200c_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    for (auto& e : c->ev_stage) cudaEventCreate(&e);
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (auto& e : c->ev_in) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    DevTables* h = new DevTables(); build_tables(h);
    if (cudaMalloc(&c->d_tables, sizeof(DevTables)) != cudaSuccess) { delete h; delete c; return nullptr; }
    cudaMemcpy(c->d_tables, h, sizeof(DevTables), cudaMemcpyHostToDevice);
    delete h;
    c->h_pinned_cap = 1 << 16;
    if (cudaMallocHost(&c->h_pinned, c->h_pinned_cap) != cudaSuccess) { cudaFree(c->d_tables); delete c; return nullptr; }
    cudaFuncSetAttribute(k_compress_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 65536 + 16);
    (void)workspace_bytes;
    return c;
}

void b200c_destroy(b200c_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& b : c->ws) if (b.p) cudaFree(b.p);
    if (c->d_tables) cudaFree(c->d_tables);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    for (auto& e : c->ev_stage) cudaEventDestroy(e);
    for (auto& e : c->ev_in) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

const char* b200c_last_error(b200c_ctx* c) { return c ? c->err.c_str() : "no context (no CUDA device?)"; }

int b200c_host_register(void* p, size_t n) { cudaError_t e = cudaHostRegister(p, n, cudaHostRegisterDefault); if (e != cudaSuccess) { cudaGetLastError(); return B200C_ECUDA; } return B200C_OK; }
int b200c_host_unregister(void* p) { cudaError_t e = cudaHostUnregister(p); if (e != cudaSuccess) { cudaGetLastError(); return B200C_ECUDA; } return B200C_OK; }

int b200c_dev_alloc(b200c_ctx* c, size_t n, void** dptr) {
    if (!c || !dptr) return B200C_EINVAL;
    cudaSetDevice(c->device);
    B200C_CUDA_TRY(c, cudaMalloc(dptr, n + 256));
    return B200C_OK;
}
int b200c_dev_free(b200c_ctx* c, void* dptr) { if (!c) return B200C_EINVAL; cudaSetDevice(c->device); B200C_CUDA_TRY(c, cudaFree(dptr)); return B200C_OK; }
int b200c_memcpy_h2d(b200c_ctx* c, void* d, const void* s, size_t n) {
    if (!c) return B200C_EINVAL;
    B200C_CUDA_TRY(c, cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, c->stream));
    B200C_CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return B200C_OK;
}
int b200c_memcpy_d2h(b200c_ctx* c, void* d, const void* s, size_t n) {
    if (!c) return B200C_EINVAL;
    B200C_CUDA_TRY(c, cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, c->stream));
    B200C_CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return B200C_OK;
}
int b200c_sync(b200c_ctx* c) { if (!c) return B200C_EINVAL; B200C_CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return B200C_OK; }
double b200c_last_kernel_ms(b200c_ctx* c) { return c ? c->last_ms : 0.0; }
uint64_t b200c_last_kernel_launches(b200c_ctx* c) { return c ? c->launches_call : 0; }
uint64_t b200c_total_kernel_launches(b200c_ctx* c) { return c ? c->launches_total : 0; }
int b200c_last_stage_ms(b200c_ctx* c, double* out, int n) {
    if (!c || !out) return 0;
    int k = c->nstages < n ? c->nstages : n;
    for (int i = 0; i < k; i++) out[i] = c->stage_ms[i];
    return k;
}

uint64_t b200c_chunk_count(uint64_t n, int chunk_len) { return chunk_len > 0 ? (n + chunk_len - 1) / chunk_len : 0; }
uint64_t b200c_compress_bound(int comp, uint64_t n, int chunk_len) {
    uint64_t nch = b200c_chunk_count(n, chunk_len);
    int m = chunk_max_compressed(comp, chunk_len); if (m < chunk_len) m = chunk_len;
    return nch * ((uint64_t)m + 4) + 64;
}
int b200c_initial_compressed_buffer_length(int comp, int chunk_len) { return chunk_max_compressed(comp, chunk_len); }

static int check_codec_args(b200c_ctx* c, int comp, int chunk_len) {
    if (!c) return B200C_EINVAL;
    if (comp != COMP_LZ4 && comp != COMP_SNAPPY && comp != COMP_NONE) { c->err = "unknown compressor"; return B200C_EINVAL; }
    if (chunk_len <= 0 || chunk_len > 65536 || (chunk_len & (chunk_len - 1))) { c->err = "chunk_len must be a power of two <= 64 KiB"; return B200C_EUNSUPPORTED; }
    return B200C_OK;
}

int b200c_compress_chunks(b200c_ctx* c, int comp, const uint8_t* in, uint64_t n, int chunk_len, int max_clen,
                          uint8_t* out, uint64_t out_cap, uint64_t* out_len, uint64_t* chunk_offsets, uint32_t* digest, int flags) {
    B200C_TRY(check_codec_args(c, comp, chunk_len));
    if ((!in && n) || !out || !out_len || !digest) { c->err = "null argument"; return B200C_EINVAL; }
    cudaSetDevice(c->device);
    const bool dev = flags & B200C_FLAG_DEVICE_PTRS;
    uint64_t nchunks = b200c_chunk_count(n, chunk_len);
    const uint8_t* d_in = in; uint8_t* d_out = out; uint64_t* d_offs;
    if (!dev) {
        uint8_t* p; B200C_TRY(ws_typed(c, WSC_IN, n + 64, &p));
        B200C_CUDA_TRY(c, cudaMemcpyAsync(p, in, n, cudaMemcpyHostToDevice, c->stream)); d_in = p;
        B200C_TRY(ws_typed(c, WSC_OUT, out_cap, &d_out));
    }
    if (dev && chunk_offsets) d_offs = chunk_offsets;     // caller provides nchunks+1 entries on the device
    else B200C_TRY(ws_typed(c, WSC_OFFS, nchunks + 1, &d_offs));
    timing_begin(c);
    int rc = compress_stream_device(c, comp, d_in, n, chunk_len, max_clen, d_out, out_cap, d_offs, out_len, digest, 0);
    int rc2 = timing_end(c);
    if (rc != B200C_OK) return rc;
    if (rc2 != B200C_OK) return rc2;
    if (!dev) {
        B200C_CUDA_TRY(c, cudaMemcpyAsync(out, d_out, *out_len, cudaMemcpyDeviceToHost, c->stream));
        if (chunk_offsets && nchunks) B200C_CUDA_TRY(c, cudaMemcpyAsync(chunk_offsets, d_offs, nchunks * 8, cudaMemcpyDeviceToHost, c->stream));
        B200C_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return B200C_OK;
}

int b200c_decompress_chunks(b200c_ctx* c, int comp, const uint8_t* data, uint64_t data_len, const uint64_t* chunk_offsets, uint64_t nchunks,
                            int chunk_len, int max_clen, uint64_t data_length, uint8_t* out, int verify_crc, b200c_corruption* where, int flags) {
    B200C_TRY(check_codec_args(c, comp, chunk_len));
    if ((!data && data_len) || (!chunk_offsets && nchunks) || (!out && data_length)) { c->err = "null argument"; return B200C_EINVAL; }
    if (nchunks != b200c_chunk_count(data_length, chunk_len)) { c->err = "chunk count does not match data_length"; return B200C_EINVAL; }
    cudaSetDevice(c->device);
    const bool dev = flags & B200C_FLAG_DEVICE_PTRS;
    const uint8_t* d_data = data; const uint64_t* d_offs = chunk_offsets; uint8_t* d_out = out;
    if (!dev) {
        uint8_t* p; B200C_TRY(ws_typed(c, WSC_IN, data_len + 64, &p));
        B200C_CUDA_TRY(c, cudaMemcpyAsync(p, data, data_len, cudaMemcpyHostToDevice, c->stream)); d_data = p;
        uint64_t* po; B200C_TRY(ws_typed(c, WSC_CHOFFS, nchunks + 1, &po));
        B200C_CUDA_TRY(c, cudaMemcpyAsync(po, chunk_offsets, nchunks * 8, cudaMemcpyHostToDevice, c->stream)); d_offs = po;
        B200C_TRY(ws_typed(c, WSC_OUT, data_length + 64, &d_out));
    }
    ChunkErr* d_err; B200C_TRY(ws_typed(c, WSC_ERR, 1, &d_err));
    B200C_CUDA_TRY(c, cudaMemsetAsync(d_err, 0xFF, sizeof(ChunkErr), c->stream));
    timing_begin(c);
    int rc = decompress_stream_device(c, comp, d_data, data_len, d_offs, nchunks, chunk_len, max_clen, data_length, d_out, verify_crc, d_err);
    int rc2 = timing_end(c);
    if (rc != B200C_OK) return rc;
    if (rc2 != B200C_OK) return rc2;
    ChunkErr* h = (ChunkErr*)c->h_pinned;
    B200C_CUDA_TRY(c, cudaMemcpyAsync(h, d_err, sizeof(ChunkErr), cudaMemcpyDeviceToHost, c->stream));
    B200C_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (h->first_bad != ~0ull) {
        uint64_t chunk = h->first_bad >> 8; int kind = (int)(h->first_bad & 0xff);
        if (where) { where->input = 0; where->kind = kind; where->chunk = chunk; where->offset = 0; }
        c->err = std::string(kind == 1 ? "chunk CRC mismatch" : "malformed compressed chunk") + " at chunk " + std::to_string(chunk);
        return B200C_ECORRUPT;
    }
    if (!dev && data_length) {
        B200C_CUDA_TRY(c, cudaMemcpyAsync(out, d_out, data_length, cudaMemcpyDeviceToHost, c->stream));
        B200C_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return B200C_OK;
}

// ICompressor.compress: one buffer in, compressor output (no CRC) out.
int b200c_compress(b200c_ctx* c, int comp, const uint8_t* in, int n, uint8_t* out, int out_cap) {
    if (!c || n < 0 || !out) return B200C_EINVAL;
    if (n > 65536) { c->err = "single-buffer compress is limited to 64 KiB"; return B200C_EUNSUPPORTED; }
    int chunk_len = 1; while (chunk_len < n) chunk_len <<= 1;
    if (n == 0) {   // degenerate: LZ4 of nothing = length prefix + one empty-literal token; Snappy = varint 0
        if (comp == COMP_LZ4) { if (out_cap < 5) return B200C_ETOOSMALL; memset(out, 0, 5); return 5; }
        if (comp == COMP_SNAPPY) { if (out_cap < 1) return B200C_ETOOSMALL; out[0] = 0; return 1; }
        return 0;
    }
    std::vector<uint8_t> tmp(b200c_compress_bound(comp, n, chunk_len));
    uint64_t out_len = 0, off = 0; uint32_t dig = 0;
    int rc = b200c_compress_chunks(c, comp, in, (uint64_t)n, chunk_len, INT_MAX, tmp.data(), tmp.size(), &out_len, &off, &dig, 0);
    if (rc != B200C_OK) return rc;
    int clen = (int)out_len - 4;
    if (clen > out_cap) { c->err = "output buffer too small"; return B200C_ETOOSMALL; }
    memcpy(out, tmp.data(), clen);
    return clen;
}

// ICompressor.uncompress: compressor output in, plain bytes out; returns the decoded length.
int b200c_uncompress(b200c_ctx* c, int comp, const uint8_t* in, int n, uint8_t* out, int out_cap) {
    if (!c || !in || n <= 0 || out_cap < 0) return B200C_EINVAL;
    int ulen;
    if (comp == COMP_LZ4) { if (n < 4) { c->err = "truncated"; return B200C_ECORRUPT; } ulen = (int)((uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16) | ((uint32_t)in[3] << 24)); }
    else if (comp == COMP_SNAPPY) { uint32_t v = 0; int sh = 0, i = 0; bool ok = false; for (; i < n && i < 5; i++) { v |= (uint32_t)(in[i] & 0x7f) << sh; if (!(in[i] & 0x80)) { ok = true; break; } sh += 7; } if (!ok) { c->err = "bad varint"; return B200C_ECORRUPT; } ulen = (int)v; }
    else ulen = n;
    if (ulen < 0 || ulen > 65536) { c->err = "single-buffer uncompress is limited to 64 KiB"; return B200C_ECORRUPT; }
    if (ulen > out_cap) { c->err = "output buffer too small"; return B200C_ETOOSMALL; }
    if (ulen == 0) return 0;
    int chunk_len = 1; while (chunk_len < ulen) chunk_len <<= 1;
    std::vector<uint8_t> img(n + 4); memcpy(img.data(), in, n); memset(img.data() + n, 0, 4);
    uint64_t off = 0;
    int rc = b200c_decompress_chunks(c, comp, img.data(), img.size(), &off, 1, chunk_len, INT_MAX, (uint64_t)ulen, out, 0, nullptr, 0);
    if (rc != B200C_OK) return rc;
    return ulen;
}

} // extern "C"
