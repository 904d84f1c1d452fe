// zb_checksum.cu — CRC-32 / Adler-32 kernels and their launch logic.
// See zb_checksum.cuh for the algorithm.  Two kernel shapes:
//
//   ck_big  : one (huge) segment, one CTA per SM, 1024 threads, lane-private
//             (bank-conflict-free) Horner tables in 128 KiB of shared memory.
//             HBM-read bound: algorithmic bytes = len, one pass.
//   ck_seg  : many segments (gzip members / chunks), 256 threads per CTA,
//             compact 4 KiB tables, items = (segment, part) pairs grid-strided.
//
// Both XOR / add their weighted per-CTA partials into a per-segment CkAccum;
// ck_finish_kernel folds in the caller's running value and writes the results.
#include "zb_internal.h"
#include "zb_checksum.cuh"

namespace zb {

constexpr int kBigThreads = 1024;     // x kBigUnroll 16-byte loads in flight per thread
constexpr int kBigUnroll = 4;         // (512 threads x 8 loads measured 3.7 TB/s vs 5.8: too few warps to cover LDS latency)
constexpr int kSegThreads = 256;

template <int T>
__device__ __forceinline__ void block_reduce_store(CkPartial v, CkAccum *acc, bool do_crc, bool do_adler) {
    __shared__ uint32_t red[3][T / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        v.crc ^= __shfl_xor_sync(0xffffffffu, v.crc, d);
        v.a += __shfl_xor_sync(0xffffffffu, v.a, d);
        v.b += __shfl_xor_sync(0xffffffffu, v.b, d);
    }
    if (lane == 0) { red[0][warp] = v.crc; red[1][warp] = v.a; red[2][warp] = v.b; }
    __syncthreads();
    if (warp == 0) {
        uint32_t c = lane < T / 32 ? red[0][lane] : 0;
        uint32_t a = lane < T / 32 ? red[1][lane] : 0;
        uint32_t b = lane < T / 32 ? red[2][lane] : 0;
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            c ^= __shfl_xor_sync(0xffffffffu, c, d);
            a += __shfl_xor_sync(0xffffffffu, a, d);
            b += __shfl_xor_sync(0xffffffffu, b, d);
        }
        if (lane == 0) {
            if (do_crc && c) atomicXor(&acc->crc, c);
            if (do_adler) {
                atomicAdd(&acc->a, (unsigned long long)a);
                atomicAdd(&acc->b, (unsigned long long)b);
            }
        }
    }
    __syncthreads();
}

// ---- big: one segment over the whole grid ---------------------------------
template <bool DO_CRC, bool DO_ADLER>
__global__ void __launch_bounds__(kBigThreads, 1)
ck_big_kernel(const uint8_t *__restrict__ data, uint64_t len, const CrcTables *__restrict__ tabs,
              CkAccum *__restrict__ acc) {
    extern __shared__ __align__(16) uint32_t smem[];
    // Lane-private (bank-conflict-free) copies of the four 256-entry tables.  Byte
    // offset of (table j, byte b, lane l):  (j>>1)*65536 + b*256 + (j&1)*128 + l*4,
    // i.e. tables are interleaved in pairs so that one PRMT builds `b*256 + l*4`
    // (byte b of the state into bits 8..15, the lane offset into bits 0..7) and the
    // rest of the address is an immediate of the LDS.
    uint32_t *rep = smem;
    uint32_t *x2n = smem + 4 * 256 * 32;         // [32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (DO_CRC) {
        const uint32_t *src = &tabs->big[0][0];
        for (int e = warp; e < 1024; e += kBigThreads / 32) {
            const int j = e >> 8, bb = e & 255;
            rep[(j >> 1) * 16384 + bb * 64 + (j & 1) * 32 + lane] = src[e];
        }
    }
    if (threadIdx.x < 32) x2n[threadIdx.x] = tabs->x2n[threadIdx.x];
    __syncthreads();
    const uint32_t x32 = tabs->x32;

    const CkPart part = ck_make_part(data, len, blockIdx.x, gridDim.x);
    const uint32_t lane4 = (uint32_t)lane * 4u;
    const char *repb = reinterpret_cast<const char *>(rep);
    auto tab = [repb, lane4](int j, uint32_t v) -> uint32_t {
        // selector: result byte0 <- lane4.byte0, byte1 <- v.byte j, bytes 2,3 <- lane4.byte1 (= 0)
        const uint32_t off = __byte_perm(v, lane4, 0x5504u | ((uint32_t)j << 4));
        return *reinterpret_cast<const uint32_t *>(repb + off + (j >> 1) * 65536 + (j & 1) * 128);
    };
    CkPartial v = ck_thread_body<DO_CRC, DO_ADLER, kBigUnroll>(part, threadIdx.x, kBigThreads, tab, x2n, x32);
    if (threadIdx.x == 0) {
        CkPartial e = ck_edge_bytes<DO_CRC, DO_ADLER>(part, x2n);
        v.crc ^= e.crc; v.a += e.a; v.b += e.b;
    }
    block_reduce_store<kBigThreads>(v, acc, DO_CRC, DO_ADLER);
}

// ---- seg: many segments, items grid-strided -------------------------------
template <bool DO_CRC, bool DO_ADLER>
__global__ void __launch_bounds__(kSegThreads)
ck_seg_kernel(const uint8_t *__restrict__ base, const uint64_t *__restrict__ seg_off,
              const uint64_t *__restrict__ seg_len, uint64_t single_len, uint32_t nseg, uint32_t parts,
              const CrcTables *__restrict__ tabs, CkAccum *__restrict__ acc) {
    __shared__ uint32_t tab_s[4 * 256];
    __shared__ uint32_t x2n[32];
    for (int e = threadIdx.x; e < 1024; e += kSegThreads) tab_s[e] = (&tabs->seg[0][0])[e];
    if (threadIdx.x < 32) x2n[threadIdx.x] = tabs->x2n[threadIdx.x];
    __syncthreads();
    const uint32_t x32 = tabs->x32;
    auto tab = [&](int j, uint32_t v) -> uint32_t { return tab_s[j * 256 + ((v >> (8 * j)) & 0xff)]; };

    const uint64_t items = (uint64_t)nseg * parts;
    for (uint64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const uint32_t seg = (uint32_t)(it / parts), pi = (uint32_t)(it % parts);
        const uint64_t off = seg_off ? seg_off[seg] : 0;
        const uint64_t len = seg_len ? seg_len[seg] : single_len;
        const CkPart part = ck_make_part(base + off, len, pi, parts);
        CkPartial v = ck_thread_body<DO_CRC, DO_ADLER>(part, threadIdx.x, kSegThreads, tab, x2n, x32);
        if (threadIdx.x == 0) {
            CkPartial e = ck_edge_bytes<DO_CRC, DO_ADLER>(part, x2n);
            v.crc ^= e.crc; v.a += e.a; v.b += e.b;
        }
        block_reduce_store<kSegThreads>(v, acc + seg, DO_CRC, DO_ADLER);
    }
}

__global__ void ck_finish_kernel(CkAccum *__restrict__ acc, const uint64_t *__restrict__ seg_len,
                                 uint64_t single_len, uint32_t nseg, int which, uint32_t init_crc,
                                 uint32_t init_adler, const uint32_t *d_init2, const CrcTables *__restrict__ tabs,
                                 uint32_t *d_crc, uint32_t *d_adler) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    const uint64_t len = seg_len ? seg_len[i] : single_len;
    if (d_init2) { init_crc = d_init2[0]; init_adler = d_init2[1]; }   // running values chained on the device
    uint32_t crc, adler;
    ck_finish(acc[i].crc, acc[i].a, acc[i].b, len, init_crc, init_adler, tabs->x2n, &crc, &adler);
    if ((which & ZB200_CRC32) && d_crc) d_crc[i] = crc;
    if ((which & ZB200_ADLER32) && d_adler) d_adler[i] = adler;
}

// ---------------------------------------------------------------------------
int checksum_init(zb200_ctx *ctx) {
    CrcTables h;
    memset(&h, 0, sizeof h);
    X2nTable x2n;
    gf2_fill_x2n(x2n);
    for (int k = 0; k < 32; ++k) h.x2n[k] = x2n.v[k];
    h.x32 = x2n.v[5];
    ck_fill_horner(h.big, x2n.v, kBigThreads);
    ck_fill_horner(h.seg, x2n.v, kSegThreads);
    ZB_CUDA(cudaMalloc(&ctx->d_crc_tables, sizeof h));
    ZB_CUDA(cudaMemcpy(ctx->d_crc_tables, &h, sizeof h, cudaMemcpyHostToDevice));
    const int big_smem = (4 * 256 * 32 + 32) * 4;
    ZB_CUDA(cudaFuncSetAttribute(ck_big_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    ZB_CUDA(cudaFuncSetAttribute(ck_big_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    ZB_CUDA(cudaFuncSetAttribute(ck_big_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem));
    return ZB200_OK;
}

// Bytes below which a single segment is cheaper through the small-CTA kernel.
constexpr uint64_t kBigMin = 8ull << 20;

int checksum_launch(zb200_ctx *ctx, const uint8_t *d_base, const uint64_t *d_off, const uint64_t *d_len,
                    uint64_t single_len, size_t nseg, int which, uint32_t init_crc, uint32_t init_adler,
                    uint32_t *d_crc, uint32_t *d_adler, CkAccum *d_acc, cudaStream_t s, const uint32_t *d_init2) {
    if (nseg == 0) return ZB200_OK;
    if (!(which & 3) || nseg > 0xffffffffull) return ZB200_ERR_PARAM;
    const bool crc = which & ZB200_CRC32, adl = which & ZB200_ADLER32;
    ZB_CUDA(cudaMemsetAsync(d_acc, 0, nseg * sizeof(CkAccum), s));
    if (nseg == 1 && !d_off && single_len >= kBigMin) {
        const int big_smem = (4 * 256 * 32 + 32) * 4;
        const int grid = ctx->sm_count;
        prof_mark(ctx, s, "ck_big_kernel");
        if (crc && adl) ck_big_kernel<true, true><<<grid, kBigThreads, big_smem, s>>>(d_base, single_len, ctx->d_crc_tables, d_acc);
        else if (crc)   ck_big_kernel<true, false><<<grid, kBigThreads, big_smem, s>>>(d_base, single_len, ctx->d_crc_tables, d_acc);
        else            ck_big_kernel<false, true><<<grid, kBigThreads, big_smem, s>>>(d_base, single_len, ctx->d_crc_tables, d_acc);
    } else {
        // parts per segment: single segment -> spread over the GPU in >= 64 KiB parts;
        // many segments -> one part each (members are 64 KiB..1 MiB).
        uint32_t parts = 1;
        if (nseg == 1 && !d_off) {
            uint64_t want = (single_len + 65535) / 65536;
            uint64_t cap = (uint64_t)ctx->sm_count * 8;
            parts = (uint32_t)(want < 1 ? 1 : want > cap ? cap : want);
        }
        uint64_t items = (uint64_t)nseg * parts;
        uint64_t cap = (uint64_t)ctx->sm_count * 8;
        const int grid = (int)(items < cap ? items : cap);
        prof_mark(ctx, s, "ck_seg_kernel");
        if (crc && adl) ck_seg_kernel<true, true><<<grid, kSegThreads, 0, s>>>(d_base, d_off, d_len, single_len, (uint32_t)nseg, parts, ctx->d_crc_tables, d_acc);
        else if (crc)   ck_seg_kernel<true, false><<<grid, kSegThreads, 0, s>>>(d_base, d_off, d_len, single_len, (uint32_t)nseg, parts, ctx->d_crc_tables, d_acc);
        else            ck_seg_kernel<false, true><<<grid, kSegThreads, 0, s>>>(d_base, d_off, d_len, single_len, (uint32_t)nseg, parts, ctx->d_crc_tables, d_acc);
    }
    ZB_LAUNCHED();
    ZB_CHECK_LAUNCH();
    const int fb = 128;
    prof_mark(ctx, s, "ck_finish_kernel");
    ck_finish_kernel<<<(unsigned)((nseg + fb - 1) / fb), fb, 0, s>>>(d_acc, d_len, single_len, (uint32_t)nseg, which,
                                                                    init_crc, init_adler, d_init2, ctx->d_crc_tables, d_crc, d_adler);
    ZB_LAUNCHED();
    ZB_CHECK_LAUNCH();
    prof_mark(ctx, s, nullptr);
    return ZB200_OK;
}

}  // namespace zb
