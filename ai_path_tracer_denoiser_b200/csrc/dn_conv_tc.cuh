// Tensor-core conv engine of the denoiser: halo-resident implicit GEMM on tcgen05 (sm_100a).
//
// One conv layer = GEMM  D[pixel, cout] = sum over (tap, cin) A[pixel + tap offset, cin] * B[cout, (tap, cin)]:
//   * M tile  = 128 output pixels = an 8-wide x 16-tall patch (pixel = TMEM lane, m = y * 8 + x),
//   * N       = padded Cout (16 .. 112, one tcgen05.mma N),
//   * K loop  = 16-channel chunks; per chunk ONE pipeline stage brings the patch's 10 x 18 halo tile of 4 channel quads
//               (11.25 KB) into shared memory with one TMA box; the 9 taps of the 3x3 stencil are then 9 x 2
//               tcgen05.mma.kind::tf32 (K = 8 each) whose A descriptors merely START at different pixels of that tile.
//               This works because activations live in HBM as channel-quad planes [C/4][rows][W][4] ("CHW4"): the TMA box
//               lands as [quad][row][pixel][16 B], which is exactly the canonical no-swizzle K-major UMMA layout -
//               8 pixels x 16 B = one 128-byte core matrix, SBO = tile row pitch (160 B), LBO = quad plane pitch (2880 B) -
//               and a no-swizzle descriptor may start at any 16-byte boundary, so a tap shift (dy, dx) is +dy*160 + dx*16 B.
//               Each activation is therefore read from L2 1.4x (halo overlap) instead of 9x; the image border (conv
//               padding = 1) is TMA out-of-bounds zero fill in x and a zero apron row above / below every tensor in y
//               (the same apron rows are the halo slots of the multi-GPU row-strip mode).
//   * B       = weights pre-packed on the host into the same core-matrix form; when a layer's whole weight set fits in
//               shared memory next to the A stages it is loaded ONCE per CTA (cp.async.bulk) and stays resident, otherwise
//               the chunk's 9 taps stream in with the A tile;
//   * concat (model.py:67,136-140) = the K loop walks two tensor maps; nearest upsample x2 (model.py:40) = four 2x2-tap
//     phase GEMMs on the half-resolution sources with the 3x3 weights pre-summed per phase (2.25x fewer MACs);
//   * epilogue (4 warps, thread = pixel): tcgen05.ld 32x32b.x16 -> bias/BatchNorm(eval)/LeakyReLU -> one float4 per channel
//     quad (a warp stores 4 x 128 contiguous bytes), MaxPool2d(2) fused through two warp shuffles (the 2x2 window lives in
//     one warp), activations rounded to tf32 (RN) so the next layer's operand truncation is exact.
// Warp roles: 0..7 = epilogue (two warps per TMEM lane quarter, alternating 16-column chunks, so every SM sub-partition has two
// epilogue warps to hide latency), 8 = TMA producer, 9 = TMEM allocator + barrier init, 10 and 11 = MMA issuers (one lane each) that take the
// CTA's tiles alternately.  Pipelines: up
// to 8 smem stages (full/empty mbarriers) and 2 TMEM accumulator buffers (tmem_full/tmem_empty), persistent CTAs, one per SM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#include "ptd_internal.h"

#define TC_TILE_W 8
#define TC_TILE_H 16
#define TC_HALO_W (TC_TILE_W + 2)
#define TC_HALO_H (TC_TILE_H + 2)
#define TC_ROW_PITCH (TC_HALO_W * 16)                   // 160 B: one halo row of one channel quad
#define TC_QUAD_PITCH (TC_HALO_H * TC_ROW_PITCH)        // 2880 B: one channel quad of the halo tile
#define TC_A_BYTES (4 * TC_QUAD_PITCH)                  // 11520 B: 16 channels
#define TC_MAX_STAGES 8
#define TC_THREADS 384                                  // warps 0-7: epilogue; 8: TMA producer; 9: TMEM allocator + barrier init; 10, 11: MMA issuers
#define TC_EPI_WARPS 8
#define TC_WARP_TMA 8
#define TC_WARP_ALLOC 9
#define TC_WARP_INIT 9
#define TC_WARP_MMA 10                                  // first of the TC_MMA_WARPS issuing warps
#define TC_MMA_WARPS 2
#define TC_TMEM_COLS 512                                // the whole tensor memory of the SM (one CTA per SM)
#define TC_ACC_COLS 128                                 // widest accumulator (padded Cout)
#define TC_MAX_BUFS 4                                   // accumulator buffers in flight between the MMA issuer and the epilogue
#define TC_MAX_SEGS 4                                   // split-operand modes: main accumulators per buffer
#define TC_SMEM_BUDGET (200 * 1024)

// An activation tensor in HBM: channel VECTORS of 16 bytes - [cp*esize/16][rows + 2][W][16 B] - i.e. quads of fp32 ("CHW4", esize 4)
// or octets of fp16 ("CHW8", esize 2); image row r at buffer row r + 1 (rows 0 and rows + 1 are aprons).  Both element types share
// the byte geometry (one pixel of one vector = 16 B), so addresses are computed in float units for both: vector v, pixel (y, x)
// lives at base + v * quad_stride() + ((y + 1) * W + x) * 4.
struct DnTensor {
    float* base = nullptr;
    int cp = 0, rows = 0, W = 0, esize = 4;
    size_t lo_off = 0;                   // PTD_DN_3XTF32: floats from the tf32-rounded "hi" copy to the residual "lo" copy (0 = none)
    __host__ __device__ int nvec() const { return cp * esize / 16; }
    __host__ __device__ size_t quad_stride() const { return (size_t)(rows + 2) * W * 4; }      // floats (= 16-byte units * 4) between vectors
    __host__ __device__ size_t floats() const { return (size_t)nvec() * quad_stride(); }
};

struct TcConvDesc {
    DnTensor src0, src1;                 // src1.base == nullptr: single source
    int upsample;                        // sources live at half the output resolution
    DnTensor out, pool_out;              // pool_out.base == nullptr: no fused MaxPool
    bool lrelu_first;
    const float* scale; const float* shift; const float* bias;
    bool round_out = true;               // round stored activations to tf32 (all layers but the last)
    const float* shared_wpack = nullptr; // reuse another plan's packed weights (same layer, other hidden-state parity)
    int src_yoff = 0;                    // row-strip mode: first source row of this strip's tile domain inside the (replicated, full-height) sources
};

// Row-strip (multi-GPU) coupling of one conv launch, see ptd_dn.cu "row strips".  All pointers may be null (single GPU).
struct TcStripLink {
    DnTensor out_up, out_down;           // the neighbour strips' copies of `out` (peer memory): our first / last row -> their apron
    DnTensor pool_up, pool_down;
    uint32_t* sig[4];                    // flags in the neighbours' memory: out->up, out->down, pool->up, pool->down
    const uint32_t* wait[4];             // local flags: src0 from up, src0 from down, src1 from up, src1 from down
    uint32_t wait_epoch[4];
    uint32_t* done;                      // local CTA-completion counter of this layer (monotonic)
    uint32_t epoch;                      // frame sequence number written to the flags
    // gather (the level where tiling ends, ptd_dn.cu "replicated levels"): the pooled output is a full-height tensor that EVERY strip
    // holds; this strip's pooled rows (from row pool_yoff on) are stored into all of them
    float* gather_base[8];               // the other strips' copies of pool_out (peer memory); null = none / self
    uint32_t* gather_sig[8];             // their "rows of strip r have arrived" flags
    const uint32_t* gather_wait[8];      // local flags a CONSUMER of a gathered source waits for (all other strips)
    int pool_yoff;
};

struct __align__(64) TcParams {
    CUtensorMap mapA0, mapA1;
    CUtensorMap mapA0lo, mapA1lo;        // PTD_DN_3XTF32: the residual copies of the sources
    int x3;                              // 3xTF32: every chunk runs three passes, (A hi, B hi), (A hi, B lo), (A lo, B hi)
    const float* wpack;                  // packed weights, stage-major (3xTF32: per chunk the hi block then the lo block)
    int n0, n1;                          // chunks (4 channel vectors = 16 fp32 / 32 fp16 channels) of source 0 / source 1
    int v0, v1;                          // channel vectors of source 0 / source 1 (the last chunk of a source may hold only 2)
    int half;                            // fp16 storage + kind::f16 (1) or fp32 storage + kind::tf32 (0)
    int ntaps, nphases;
    int dy[4][9], dx[4][9];
    int tiles_x, tiles_y, total_items;
    int Hs, Ws;                          // tile domain (= this strip's rows at source resolution)
    int src_yoff;                        // tile row y reads source image row y + src_yoff (replicated full-height sources)
    int out_mul;                         // 1, or 2 for the upsampling layers
    int coutp;
    int stages, resident;                // smem pipeline depth; weights resident in smem (1) or streamed per stage (0)
    uint32_t b_stage_bytes;              // ntaps * coutp * 64
    uint32_t w_total_bytes;
    const float* scale; const float* shift; const float* bias;
    int lrelu_first, round_out;
    DnTensor out, pool_out;
    TcStripLink link;
    int pdl;                             // launched with programmatic stream serialization
    // accumulators in tensor memory: `nbuf` buffers of `buf_cols` columns; a buffer = nseg main accumulators [+ one correction accumulator]
    // of coutp columns each.  Split-operand modes (x3) keep the small cross terms (hi x lo, lo x hi) out of the large hi x hi sums and cut
    // the hi x hi chain into nseg pieces - tcgen05.mma truncates on every accumulate (tools/microbench/umma_accum), so one long chain loses
    // ~10x the accuracy the operand split buys; the epilogue adds the pieces in fp32 (round to nearest).
    int nseg, nbuf, buf_cols;
    int fused;                           // split modes: hi x hi and hi x lo of a chunk in ONE MMA of N = 2 * coutp (see mma_role); else three passes
    uint32_t corr_mask;                  // bit a: accumulator a (columns [a * coutp, (a + 1) * coutp) of a buffer) holds cross terms
    uint32_t b_load_bytes;               // streamed weights: bytes of B a stage carries (b_stage_bytes, or twice that when fused)
    uint32_t seg_start;                  // bit c: chunk c starts the next main accumulator
    float corr_scale;                    // the correction accumulator's scale: 1 (3xTF32) or 2^-11 (2xF16: lo operands are stored x 2^11)
    int nissue;                          // MMA issuer warps in use (2 when the stage ring splits into two rings of >= 2 stages and nbuf is even)
    int dbg;                             // PTD_DN_DEBUG (timing experiments only, results are garbage): 1 = no epilogue stores, 2 = no TMA loads, 4 = no TMEM loads
    int linked;                          // row-strip mode: some TcStripLink pointer is set (selects the kernel variant with the peer stores)
};

struct TcConvPlan {
    int valid = 0;
    TcParams p;
    float* d_wpack = nullptr;
    int grid = 0;
    size_t smem = 0;
};

// ---- device helpers (inline PTX) -----------------------------------------------------------------------------------------
namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {          // one try_wait (the hardware suspends the thread for a bounded time)
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Every in-kernel wait is bounded: a protocol error traps (the next API call returns PTD_ERR_CUDA) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    PtdSpinGuard guard;
    while (!mbar_try(bar, parity)) guard.tick();
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
// K-major operand without swizzle: core matrix = 8 rows x 16 B (128 contiguous bytes); LBO = byte distance between the core
// matrices adjacent in K, SBO = between the 8-row groups adjacent in M / N; descriptor version 1 (sm_100), layout type 0
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address, bits [0,14)
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;         // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;         // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version
    return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {    // explicit shared-space load: through a generic pointer the compiler emits LD.E (address translation, long scoreboard)
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ bool elect_one() {            // one lane of the (converged) warp, known to the compiler as such
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
template <bool HALF>
__device__ __forceinline__ void mma_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    if (HALF)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
            "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
            "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
}  // namespace tc

namespace tc {
// 16 consecutive output channels [c0, c0 + 16) of one pixel -> the tensor's channel vectors (row: address of the pixel in vector 0)
// storage formats of an activation tensor: 0 = fp32 quads, 1 = fp16 octets, 2 = fp32 hi / lo pair (3xTF32), 3 = fp16 hi / lo pair (2xF16)
__host__ __device__ __forceinline__ int tensor_format(const DnTensor& t) { return t.esize == 4 ? (t.lo_off ? 2 : 0) : (t.lo_off ? 3 : 1); }
template <int FMT>
__device__ __forceinline__ void store16_fmt(const DnTensor& t, float* row, int c0, const float* o) {
    if (FMT == 2) {                                        // 3xTF32: hi = tf32(v), lo = tf32(v - hi)
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { hi[e] = round_tf32(o[4 * qd + e]); lo[e] = round_tf32(o[4 * qd + e] - hi[e]); }   // RN, not the MMA's truncation
            float* d = row + (size_t)(c0 / 4 + qd) * t.quad_stride();
            *reinterpret_cast<float4*>(d) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(d + t.lo_off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    } else if (FMT == 3) {                                 // 2xF16: hi = f16(v), lo = f16((v - hi) * 2^11) - the residual scaled into fp16's normal range
#pragma unroll
        for (int oc = 0; oc < 2; ++oc) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float v0 = o[8 * oc + 2 * e], v1 = o[8 * oc + 2 * e + 1];
                const __half2 h = __floats2half2_rn(v0, v1);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn((v0 - hf.x) * 2048.0f, (v1 - hf.y) * 2048.0f);
                hw[e] = *reinterpret_cast<const uint32_t*>(&h); lw[e] = *reinterpret_cast<const uint32_t*>(&l);
            }
            float* d = row + (size_t)(c0 / 8 + oc) * t.quad_stride();
            *reinterpret_cast<uint4*>(d) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(d + t.lo_off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
    } else if (FMT == 0) {
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
            *reinterpret_cast<float4*>(row + (size_t)(c0 / 4 + qd) * t.quad_stride()) = make_float4(o[4 * qd], o[4 * qd + 1], o[4 * qd + 2], o[4 * qd + 3]);
    } else {
#pragma unroll
        for (int oc = 0; oc < 2; ++oc) {
            uint4 v;
            __half2 h0 = __floats2half2_rn(o[8 * oc], o[8 * oc + 1]), h1 = __floats2half2_rn(o[8 * oc + 2], o[8 * oc + 3]);
            __half2 h2 = __floats2half2_rn(o[8 * oc + 4], o[8 * oc + 5]), h3 = __floats2half2_rn(o[8 * oc + 6], o[8 * oc + 7]);
            v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1); v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(row + (size_t)(c0 / 8 + oc) * t.quad_stride()) = v;
        }
    }
}
__device__ __forceinline__ void store16(const DnTensor& t, float* row, int c0, const float* o) {      // format decided at run time (pack_gbuffer)
    switch (tensor_format(t)) {
        case 0: store16_fmt<0>(t, row, c0, o); break;
        case 1: store16_fmt<1>(t, row, c0, o); break;
        case 2: store16_fmt<2>(t, row, c0, o); break;
        default: store16_fmt<3>(t, row, c0, o); break;
    }
}
// K-major no-swizzle descriptors, split into 32-bit words: lo = start >> 4 | (LBO >> 4) << 16, hi = SBO >> 4 | version 1 << 14.
// One chunk (one shared-memory stage): NTAPS taps x (TWO ? 2 : 1) K steps into one accumulator.  Every descriptor is the stage's base plus
// a compile-time constant (padded Cout, tap geometry and K-step pitch are template parameters), so an MMA costs two uniform adds and
// the UTCHMMA itself.  a_base already carries the phase's tap origin and the LBO field, b_cur the LBO field.
// NCOLS = the MMA's N (accumulator columns written), BROWS = rows of the weight slab per k vector (its LBO / 16): coutp for both, except in
// the fused split scheme, where a slab holds the hi rows followed by the lo rows (BROWS = 2 * coutp) and one MMA of NCOLS = 2 * coutp
// computes hi x hi | hi x lo at once.
template <int NTAPS, bool HALF, int NCOLS, int BROWS, bool TWO>
__device__ __forceinline__ void issue_chunk(uint32_t d_tmem, uint32_t a_base, uint32_t b_cur, uint32_t accumulate) {
    constexpr uint32_t fmt = HALF ? 0u : 2u;                                      // D = f32, A = B = tf32 (format 2) or f16 (format 0), both K-major, M = 128
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NCOLS >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t a_hi = (uint32_t)(TC_ROW_PITCH >> 4) | (1u << 14);          // SBO = halo row pitch
    constexpr uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);                  // SBO = next 8 output channels
    constexpr uint32_t b_kstep = (uint32_t)BROWS * 2u;                            // (rows * 32 B) >> 4: one K step
#pragma unroll
    for (int t = 0; t < NTAPS; ++t) {
        // tap origin inside the halo tile in 16-byte units: 3x3 taps (dy, dx) = (t / 3 - 1, t % 3 - 1) from the tile's (1, 1); the 2x2 taps of an
        // upsampling phase (ti, tj) = (t >> 1, t & 1) from the phase's origin
        const uint32_t aoff = NTAPS == 9 ? (uint32_t)((t / 3) * (TC_ROW_PITCH >> 4) + t % 3) : (uint32_t)((t >> 1) * (TC_ROW_PITCH >> 4) + (t & 1));
        mma_w<HALF>(d_tmem, a_base + aoff, a_hi, b_cur + (uint32_t)(t * 2) * b_kstep, b_hi, idesc, t ? 1u : accumulate);
        if (TWO) mma_w<HALF>(d_tmem, a_base + aoff + (2u * TC_QUAD_PITCH >> 4), a_hi, b_cur + (uint32_t)(t * 2 + 1) * b_kstep, b_hi, idesc, 1u);
    }
}

// The issue loop is the kernel's critical path for the N = 32 layers: one M = 128 x N = 32 MMA occupies the tensor pipe's shared-memory
// operand fetch for (4096 + 1024) B / 128 B/clk = 40 cycles (tools/microbench/umma_rate, profiles/r3b_*: 41.7 cycles per MMA measured,
// whatever the layout, the data, the number of accumulators or the traffic beside it), and the MMA queue is only a couple of instructions
// deep, so whatever the issuing thread does between two MMAs is time the pipe starves (round 1: 70 cycles per tf32 MMA, 92 per f16 MMA -
// integer divisions, parameter reloads, recomputed tap offsets, ~10 instructions of descriptor arithmetic per MMA; profiles/r3c_*: a
// dozen instructions more per 18 MMAs cost 7 cycles per MMA).  Hence: nested loops instead of div/mod, every parameter hoisted into
// registers, compile-time descriptor offsets (issue_chunk), and TWO issuing warps that take the CTA's tiles alternately, so that one's
// bookkeeping (barrier polls - slow while the tensor core owns the shared-memory pipe -, loop control) hides behind the other's MMAs.
// SCHEME: 0 = one pass per chunk (tf32 / f16), 1 = split operands in three passes, 2 = split operands fused (two passes) - a template
// parameter because every instruction of the per-chunk code below is on the kernel's critical path.
template <int NTAPS, bool HALF, int COUTP, int SCHEME>
__device__ __forceinline__ void mma_role(const TcParams& p, uint8_t* smem_a, uint8_t* smem_b, uint64_t* full, uint64_t* empty,
                                         uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base, int issuer) {
    constexpr uint32_t a_lbo = (uint32_t)(TC_QUAD_PITCH >> 4) << 16;              // LBO = next channel quad
    const bool resident = p.resident != 0;
    constexpr bool x3 = SCHEME != 0, fused = SCHEME == 2;
    // fused split scheme: per chunk one stage with the hi activations - MMAs of N = 2 * coutp against slabs [hi rows | lo rows]: hi x hi into the
    // main accumulator and hi x lo into the correction accumulator right behind it, the A operand read from shared memory once for both
    // (A is 4 KB of the 5 KB an N = 32 MMA fetches) - and one stage with the lo activations: lo x hi (N = coutp, the slab's hi rows) into the
    // same correction accumulator.  Unfused: three passes hi x hi, hi x lo, lo x hi over separate hi / lo weight blocks.
    constexpr uint32_t b_lbo = (uint32_t)(fused ? 2 * COUTP : COUTP) << 16;      // (rows * 16 B) >> 4: next k vector of the same tap
    constexpr int npass = fused ? 2 : (x3 ? 3 : 1);
    const int nchunks = p.n0 + p.n1, nphases = p.nphases, total = p.total_items, stages = p.stages;
    const uint32_t sa0 = (smem_u32(smem_a) >> 4) | a_lbo, sb0 = (smem_u32(smem_b) >> 4) | b_lbo;
    const uint32_t b_stage16 = p.b_stage_bytes >> 4, a_stage16 = TC_A_BYTES >> 4, b_slot16 = p.b_load_bytes >> 4;
    const uint32_t b_chunk16 = b_stage16 * (x3 ? 2u : 1u);                    // resident weights: blocks of one chunk (hi [, lo])
    const uint32_t d_tmem0 = __shfl_sync(0xffffffffu, tmem_base, 0);          // provably warp-uniform
    // chunks whose source holds only two channel vectors there (one K step; the other two vectors are TMA zero fill): bit c
    uint32_t one_step = 0;
    for (int c = 0; c < nchunks; ++c)
        if ((c < p.n0 ? p.v0 - 4 * c : p.v1 - 4 * (c - p.n0)) <= 2) one_step |= 1u << c;
    const int nbuf = p.nbuf, nseg = p.nseg;
    const uint32_t buf_cols = (uint32_t)p.buf_cols, seg_start = p.seg_start;
    // Issuer w takes the CTA's tiles w, w + nissue, ... and owns ring w of the shared-memory stages (stages [w * S, (w + 1) * S), filled by
    // the producer with exactly this issuer's tiles) and every nissue-th accumulator buffer (nbuf is even when nissue == 2): every mbarrier
    // keeps a single waiter that sees each of its phases, so the parity waits cannot alias, no MMA of the other issuer ever touches this
    // issuer's stages or accumulators (the first MMA into an accumulator overwrites it), and each tcgen05.commit covers the issuer's own MMAs.
    const int nissue = p.nissue;
    if (issuer >= nissue) return;
    const int ring_stages = stages / nissue, stage0 = issuer * ring_stages;
    const uint32_t sa_ring = sa0 + (uint32_t)stage0 * a_stage16;
    int stage = 0; uint32_t phase = 0;                                        // position in this issuer's ring
    int buf = issuer; uint32_t buf_phase = 0;
    for (int item = blockIdx.x + issuer * (int)gridDim.x; item < total; item += nissue * (int)gridDim.x) {
        uint32_t b_base = sb0, a_ph = 0;
        if (nphases > 1) {                                                    // the four 2x2-tap phase GEMMs of an upsampling layer
            const int ph = item & 3;
            a_ph = (uint32_t)((ph >> 1) * (TC_ROW_PITCH >> 4) + (ph & 1));     // tap (0, 0) of phase (a, b) starts a rows / b pixels into the halo tile
            b_base += (uint32_t)(ph * nchunks) * b_chunk16;
        }
        uint32_t a_base = sa_ring + (uint32_t)stage * a_stage16 + a_ph;
        mbar_wait(&tmem_empty[buf], buf_phase ^ 1);
        fence_after_sync();
        const uint32_t d_buf = d_tmem0 + (uint32_t)buf * buf_cols;
        uint32_t started = 0, seg = 0;                                        // accumulators of this buffer that hold a partial sum already
        for (int c = 0; c < nchunks; ++c) {
            const bool two = !((one_step >> c) & 1u);
            if (x3 && c && ((seg_start >> c) & 1u)) ++seg;
#pragma unroll
            for (int pass = 0; pass < npass; ++pass) {
                mbar_wait(&full[stage0 + stage], phase);
                fence_after_sync();
                // resident weights: block (phase, chunk): [hi block, lo block] (passes 0 and 2 use the hi block, pass 1 the lo block) or one fused slab
                const uint32_t b_cur = resident ? b_base + ((!fused && pass == 1) ? b_stage16 : 0u) : sb0 + (uint32_t)(stage0 + stage) * b_slot16;
                const bool last = c == nchunks - 1 && pass == npass - 1;
                // accumulator: unfused - hi x hi -> main `seg`, cross terms -> the correction accumulator behind the mains; fused - pairs [main | correction]
                uint32_t d_tmem = d_buf, accumulate = c ? 1u : 0u;
                if (x3) {
                    const uint32_t ai = fused ? 2u * seg + (uint32_t)pass : (pass == 0 ? seg : (uint32_t)nseg);
                    d_tmem = d_buf + ai * (uint32_t)COUTP;
                    accumulate = (started >> ai) & 1u;
                    started |= (fused && pass == 0) ? (3u << ai) : (1u << ai);
                }
                if (elect_one()) {
                    if (fused && pass == 0) {
                        if (two) issue_chunk<NTAPS, HALF, 2 * COUTP, 2 * COUTP, true>(d_tmem, a_base, b_cur, accumulate);
                        else issue_chunk<NTAPS, HALF, 2 * COUTP, 2 * COUTP, false>(d_tmem, a_base, b_cur, accumulate);
                    } else if (fused) {
                        if (two) issue_chunk<NTAPS, HALF, COUTP, 2 * COUTP, true>(d_tmem, a_base, b_cur, accumulate);
                        else issue_chunk<NTAPS, HALF, COUTP, 2 * COUTP, false>(d_tmem, a_base, b_cur, accumulate);
                    } else {
                        if (two) issue_chunk<NTAPS, HALF, COUTP, COUTP, true>(d_tmem, a_base, b_cur, accumulate);
                        else issue_chunk<NTAPS, HALF, COUTP, COUTP, false>(d_tmem, a_base, b_cur, accumulate);
                    }
                    mma_commit(&empty[stage0 + stage]);                       // frees the smem stage when the MMAs retire
                    if (last) mma_commit(&tmem_full[buf]);                    // accumulators complete -> epilogue
                }
                __syncwarp();
                a_base += a_stage16;
                if (++stage == ring_stages) { stage = 0; phase ^= 1; a_base = sa_ring + a_ph; }
            }
            b_base += b_chunk16;
        }
        buf += nissue;
        if (buf >= nbuf) { buf -= nbuf; buf_phase ^= 1; }
    }
}
}  // namespace tc

// shared memory carve-up (offsets from the 1024-aligned base): [A stage 0..S) | B (resident: whole layer; streamed: S stages) | barriers
// FMT: storage format of the output (and pooled output) tensor - a template parameter, like LINKED, to keep the epilogue loop small: it
// shares the instruction cache with the producer and the MMA issuers.
template <bool HALF, bool LINKED, int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.stages * TC_A_BYTES;
    const size_t b_bytes = p.resident ? (size_t)p.w_total_bytes : (size_t)p.stages * p.b_load_bytes;
    uint64_t* full = (uint64_t*)(smem_b + ((b_bytes + 15) & ~(size_t)15));
    uint64_t* empty = full + TC_MAX_STAGES;
    uint64_t* tmem_full = empty + TC_MAX_STAGES;
    uint64_t* tmem_empty = tmem_full + TC_MAX_BUFS;
    uint64_t* wfull = tmem_empty + TC_MAX_BUFS;
    uint32_t* tmem_base_slot = (uint32_t*)(wfull + 1);
    float4* s_par = (float4*)(((uintptr_t)(tmem_base_slot + 1) + 15) & ~(uintptr_t)15);                  // per output channel: (s1, b1, s2, b2), out = lrelu(acc*s1 + b1)*s2 + b2

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = p.n0 + p.n1;

    // PTD_DN_PDL (opt-in): programmatic dependent launch.  The NEXT layer's CTAs may start as soon as every CTA of this launch is
    // running; they set up barriers, TMEM and their resident weights - none of which depend on this layer - and only then wait
    // (griddepcontrol.wait, TMA producer below) for this launch to finish and flush before reading its output.  The layers at
    // <= 1/8 resolution have fewer tiles than SMs and sit at a 12-17 us launch + prologue floor each; this hides the prologue.
    if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == TC_WARP_TMA && lane == 0) {
        tc::prefetch_tmap(&p.mapA0);
        if (p.n1) tc::prefetch_tmap(&p.mapA1);
        if (p.x3) { tc::prefetch_tmap(&p.mapA0lo); if (p.n1) tc::prefetch_tmap(&p.mapA1lo); }
    }
    if (warp == TC_WARP_INIT && lane == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int a = 0; a < TC_MAX_BUFS; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], TC_EPI_WARPS); }
        tc::mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_WARP_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < TC_EPI_WARPS) {
        // conv bias + BatchNorm(eval) + LeakyReLU as one branch-free form: BN->LReLU layers use (scale, shift', 1, 0),
        // the LReLU->BN layer (encoder layer2's first conv, model.py:30-32) uses (1, bias, scale, shift)
        for (int c = threadIdx.x; c < p.coutp; c += TC_EPI_WARPS * 32) {
            const float sc = __ldg(&p.scale[c]), sh = __ldg(&p.shift[c]), bi = __ldg(&p.bias[c]);
            s_par[c] = p.lrelu_first ? make_float4(1.0f, bi, sc, sh) : make_float4(sc, sh, 1.0f, 0.0f);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == TC_WARP_TMA) {
        // ===== TMA producer: warp-uniform loop, one elected lane issues (same reason as the MMA role below) ==========
        // row-strip mode: the apron rows of the sources are written by the neighbour GPUs; wait for this frame's flags
        bool waited = false;
        PtdSpinGuard guard;                                            // traps after PTD_SPIN_TIMEOUT_NS instead of hanging the GPU
        if (LINKED)
        for (int i = 0; i < 4; ++i)
            if (p.link.wait[i]) { while ((int)(tc::ld_acquire_sys(p.link.wait[i]) - p.link.wait_epoch[i]) < 0) guard.tick(); waited = true; }
        if (LINKED)
        for (int i = 0; i < 8; ++i)
            if (p.link.gather_wait[i]) { while ((int)(tc::ld_acquire_sys(p.link.gather_wait[i]) - p.link.epoch) < 0) guard.tick(); waited = true; }
        if (waited) asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire -> async-proxy (TMA) reads
        if (p.resident && tc::elect_one()) {                         // the layer's whole weight set, once per CTA
            tc::mbar_expect_tx(wfull, p.w_total_bytes);
            for (uint32_t off = 0; off < p.w_total_bytes; off += 32768u) {
                const uint32_t n = p.w_total_bytes - off < 32768u ? p.w_total_bytes - off : 32768u;
                tc::bulk_load(smem_b + off, (const uint8_t*)p.wpack + off, n, wfull);
            }
        }
        __syncwarp();
        if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");     // the previous launch has completed and its stores are visible
        const uint32_t stage_tx = TC_A_BYTES + (p.resident ? 0u : p.b_load_bytes);
        const int fused = p.fused, npass = fused ? 2 : (p.x3 ? 3 : 1), n0 = p.n0, nissue = p.nissue, ring_stages = p.stages / nissue;
        // One stage ring per MMA issuer; the CTA's tiles alternate between the issuers, and the producer feeds the `nissue` tiles in flight
        // unit by unit in turn, so that both issuers always have operands staged (feeding tile after tile would starve the second issuer
        // whenever a tile has more units than a ring has stages).
        int rstage[TC_MMA_WARPS] = {0, 0}; uint32_t rphase[TC_MMA_WARPS] = {0, 0};
        for (int item0 = blockIdx.x; item0 < p.total_items; item0 += nissue * (int)gridDim.x) {
            int x0[TC_MMA_WARPS], y0[TC_MMA_WARPS], ph[TC_MMA_WARPS]; bool live[TC_MMA_WARPS];
#pragma unroll
            for (int r = 0; r < TC_MMA_WARPS; ++r) {
                const int item = item0 + r * (int)gridDim.x;
                live[r] = r < nissue && item < p.total_items;
                ph[r] = p.nphases > 1 ? (item & 3) : 0;
                const int tile = p.nphases > 1 ? (item >> 2) : item;
                const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
                x0[r] = tx * TC_TILE_W; y0[r] = ty * TC_TILE_H + p.src_yoff;
            }
            for (int c = 0; c < nchunks; ++c) {
                for (int pass = 0; pass < npass; ++pass) {
#pragma unroll
                    for (int r = 0; r < TC_MMA_WARPS; ++r) {
                        if (!live[r]) continue;
                        const int sidx = r * ring_stages + rstage[r];
                        tc::mbar_wait(&empty[sidx], rphase[r] ^ 1);
                        if (p.dbg & 2) { if (tc::elect_one()) tc::mbar_arrive(&full[sidx]); }
                        else if (tc::elect_one()) {
                            tc::mbar_expect_tx(&full[sidx], stage_tx);
                            // halo tile origin: pixel (x0 - 1, y0 - 1) = buffer row y0 (apron offset +1); x < 0 / x >= W are zero filled
                            const bool lo = pass == (fused ? 1 : 2);                     // the pass that multiplies the lo activations
                            const CUtensorMap* map = c < n0 ? (lo ? &p.mapA0lo : &p.mapA0) : (lo ? &p.mapA1lo : &p.mapA1);
                            tc::tma_load_3d(smem_a + (size_t)sidx * TC_A_BYTES, map, &full[sidx], (x0[r] - 1) * (HALF ? 8 : 4), y0[r], (c < n0 ? c : c - n0) * 4);
                            if (!p.resident) {
                                // unfused split: blocks (hi, lo) per chunk, pass 1 takes the lo block; fused: the chunk's slab [hi rows | lo rows] for both passes
                                const size_t blk = fused ? (size_t)(ph[r] * nchunks + c) * 2 : (p.x3 ? (size_t)((ph[r] * nchunks + c) * 2 + (pass == 1 ? 1 : 0)) : (size_t)(ph[r] * nchunks + c));
                                tc::bulk_load(smem_b + (size_t)sidx * p.b_load_bytes, (const uint8_t*)p.wpack + blk * p.b_stage_bytes, p.b_load_bytes, &full[sidx]);
                            }
                        }
                        __syncwarp();
                        if (++rstage[r] == ring_stages) { rstage[r] = 0; rphase[r] ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= TC_WARP_MMA) {
        // ===== MMA issuer: the whole warp runs the (warp-uniform) loop so that descriptors live in uniform registers;
        //       one elected lane issues.  One tcgen05.mma costs a handful of uniform-datapath adds here - issued from divergent
        //       code the same loop cost ~135 cycles per MMA (R2UR + ELECT sequences) and was the kernel's bottleneck.
        if (p.resident) { tc::mbar_wait(wfull, 0); tc::fence_after_sync(); }
        // the scheme follows from the kernel variant's output format except for the last layer (fp32 output from split sources): SPLIT says
        // whether this variant can see split sources at all, which keeps the plain variants free of the split roles' code
#define TC_ROLE2(C, S) if (p.ntaps == 9) tc::mma_role<9, HALF, C, S>(p, smem_a, smem_b, full, empty, tmem_full, tmem_empty, tmem_base, warp - TC_WARP_MMA); \
                       else tc::mma_role<4, HALF, C, S>(p, smem_a, smem_b, full, empty, tmem_full, tmem_empty, tmem_base, warp - TC_WARP_MMA);
#define TC_ROLE(C) case C: if (!p.x3) { TC_ROLE2(C, 0) } else if (p.fused) { TC_ROLE2(C, 2) } else { TC_ROLE2(C, 1) } break;
        switch (p.coutp) { TC_ROLE(16) TC_ROLE(32) TC_ROLE(48) TC_ROLE(64) TC_ROLE(80) TC_ROLE(96) TC_ROLE(112) TC_ROLE(128) default: break; }
#undef TC_ROLE
#undef TC_ROLE2
    } else if (warp < TC_EPI_WARPS) {
        // ===== epilogue: TMEM -> registers -> bias/BN/LeakyReLU -> CHW4 ============================================
        // Two warps per SM sub-partition run this loop; with the N = 32 layers' MMAs at ~750 (f16) .. 1500 (tf32) cycles per tile it has
        // to stay at a few hundred instructions per tile: the row-strip stores exist only in the LINKED variant of the kernel, every
        // parameter is in a register, one integer division per tile.
        const int q = warp & 3;                                    // TMEM lane quarter this warp may read
        const int half = warp >> 2;                                // which of the alternating 16-column chunks this warp takes
        const int m = q * 32 + lane;                               // pixel of the tile == TMEM lane
        const int ty = m >> 3, tx = m & 7;
        const int coutp = p.coutp, nphases = p.nphases, tiles_x = p.tiles_x, Ws = p.Ws, Hs = p.Hs, out_mul = p.out_mul, total = p.total_items;
        const int nbuf = p.nbuf, nacc = p.buf_cols / p.coutp;
        const uint32_t corr_mask = p.corr_mask;
        const uint32_t buf_cols = (uint32_t)p.buf_cols;
        const float corr_scale = p.corr_scale;
        const bool round_out = !HALF && p.round_out && !p.x3;     // fp16 storage rounds in the conversion, the split modes split in store16
        const DnTensor out = p.out, pool_out = p.pool_out;
        const uint32_t s_par_addr = tc::smem_u32(s_par);
        // While the MMAs stream operands the tensor core owns the shared-memory pipe (tools/microbench/umma_rate: an LDS.128 beside a
        // saturated N = 32 MMA stream takes ~38 cycles), so the level-0 / level-1 epilogues must not touch shared memory per tile.
        const bool reg_par = coutp <= 32, lrelu_first = p.lrelu_first != 0;
        float pa[16], pb[16], pc[16];                              // BN -> LReLU: (scale, shift', -); LReLU -> BN: (bias, scale, shift)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int c = half * 16 + j;
            const bool ok = reg_par && c < coutp;
            const float sc = ok ? __ldg(&p.scale[c]) : 0.f, sh = ok ? __ldg(&p.shift[c]) : 0.f, bi = ok ? __ldg(&p.bias[c]) : 0.f;
            pa[j] = lrelu_first ? bi : sc; pb[j] = lrelu_first ? sc : sh; pc[j] = sh;
        }
        const int pool_yoff = LINKED ? p.link.pool_yoff : 0;
        int buf = 0; uint32_t buf_phase = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const int ph = nphases > 1 ? (item & 3) : 0, tile = nphases > 1 ? (item >> 2) : item;
            const int tyi = tile / tiles_x, txi = tile - tyi * tiles_x;
            const int x = txi * TC_TILE_W + tx, y = tyi * TC_TILE_H + ty;
            const bool valid = x < Ws && y < Hs;
            const int oy = out_mul * y + (ph >> 1), ox = out_mul * x + (ph & 1);
            float* orow = out.base + ((size_t)(oy + 1) * out.W + ox) * 4;
            const size_t poff = ((size_t)((y >> 1) + pool_yoff + 1) * pool_out.W + (x >> 1)) * 4;
            float* prow = pool_out.base ? pool_out.base + poff : nullptr;
            tc::mbar_wait(&tmem_full[buf], buf_phase);
            tc::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * buf_cols;
            for (int c0 = half * 16; c0 < coutp; c0 += 32) {
                float acc[16];
                {
                    // the buffer's accumulators, summed in fp32 (round to nearest): the cross-term accumulators first (scaled), then the main segments
                    uint32_t r[16];
                    if (nacc == 1) {
                        if (p.dbg & 4) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) r[j] = (uint32_t)(j + c0);
                        } else {
                            tc::tmem_ld16(taddr + (uint32_t)c0, r);
                            tc::tmem_ld_wait();
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
                        for (int a = 0; a < nacc; ++a) {
                            if (!((corr_mask >> a) & 1u)) continue;
                            tc::tmem_ld16(taddr + (uint32_t)(a * coutp + c0), r);
                            tc::tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) acc[j] = fmaf(__uint_as_float(r[j]), corr_scale, acc[j]);
                        }
                        for (int a = nacc - 1; a >= 0; --a) {
                            if ((corr_mask >> a) & 1u) continue;
                            tc::tmem_ld16(taddr + (uint32_t)(a * coutp + c0), r);
                            tc::tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(r[j]);
                        }
                    }
                }
                float o[16];
                if (reg_par) {                                     // N <= 32: this warp's 16 channels never change - parameters live in registers
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float v = lrelu_first ? acc[j] + pa[j] : fmaf(acc[j], pa[j], pb[j]);
                        v = fmaxf(v, 0.1f * v);                    // LeakyReLU(0.1)
                        if (lrelu_first) v = fmaf(v, pb[j], pc[j]);
                        o[j] = round_out ? tc::round_tf32(v) : v;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 k = tc::lds128(s_par_addr + (uint32_t)(c0 + j) * 16u);    // broadcast LDS.128
                        float v = fmaf(acc[j], k.x, k.y);
                        v = fmaxf(v, 0.1f * v);
                        v = fmaf(v, k.z, k.w);
                        o[j] = round_out ? tc::round_tf32(v) : v;
                    }
                }
                if (valid && !(p.dbg & 1)) {
                    tc::store16_fmt<FMT>(out, orow, c0, o);
                    if (LINKED) {                                  // row-strip mode: our first / last row is the neighbour's bottom / top apron row - stored straight over NVLink
                        if (oy == 0 && p.link.out_up.base)
                            tc::store16_fmt<FMT>(p.link.out_up, p.link.out_up.base + ((size_t)(p.link.out_up.rows + 1) * out.W + ox) * 4, c0, o);
                        if (oy == out.rows - 1 && p.link.out_down.base)
                            tc::store16_fmt<FMT>(p.link.out_down, p.link.out_down.base + (size_t)ox * 4, c0, o);
                    }
                }
                if (prow) {                                        // MaxPool2d(2): partners are lanes ^1 (x) and ^8 (y)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float v = fmaxf(o[j], __shfl_xor_sync(0xffffffffu, o[j], 1));
                        o[j] = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
                    }
                    if (valid && !(tx & 1) && !(ty & 1) && !(p.dbg & 1)) {
                        tc::store16_fmt<FMT>(pool_out, prow, c0, o);
                        if (LINKED) {
#pragma unroll 1
                            for (int r = 0; r < 8; ++r)
                                if (p.link.gather_base[r]) { DnTensor t = pool_out; t.base = p.link.gather_base[r]; tc::store16_fmt<FMT>(t, t.base + poff, c0, o); }
                            if ((y >> 1) == 0 && p.link.pool_up.base)
                                tc::store16_fmt<FMT>(p.link.pool_up, p.link.pool_up.base + ((size_t)(p.link.pool_up.rows + 1) * pool_out.W + (x >> 1)) * 4, c0, o);
                            if ((y >> 1) == pool_out.rows - 1 && p.link.pool_down.base)
                                tc::store16_fmt<FMT>(p.link.pool_down, p.link.pool_down.base + (size_t)(x >> 1) * 4, c0, o);
                        }
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tmem_empty[buf]);
            if (++buf == nbuf) { buf = 0; buf_phase ^= 1; }
        }
    }
    if (LINKED && p.link.done && warp < TC_EPI_WARPS) __threadfence_system();          // our apron stores into the neighbours' memory, before the flag
    tc::fence_before_sync();
    __syncthreads();
    if (LINKED && p.link.done && threadIdx.x == 0) {
        // last CTA of the launch: every CTA's stores are ordered before its counter increment, so the flags can be raised
        __threadfence();
        const uint32_t old = atomicAdd(p.link.done, 1u);
        if (old + 1u == gridDim.x) {                                     // (the counter restarts for this layer's next launch, whatever its grid will be)
            *p.link.done = 0u;
            __threadfence_system();
            for (int i = 0; i < 4; ++i)
                if (p.link.sig[i]) tc::st_release_sys(p.link.sig[i], p.link.epoch);
            for (int i = 0; i < 8; ++i)
                if (p.link.gather_sig[i]) tc::st_release_sys(p.link.gather_sig[i], p.link.epoch);
        }
    }
    if (warp == TC_WARP_ALLOC) {
        tc::fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled tc_get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static float host_round_tf32(float x) {            // round-to-nearest (ties away, like cvt.rna) to a 10-bit mantissa
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u += 0x1000u;
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

// tensor map over a CHW4 activation: dims (W * 4 floats, rows + 2, quads); box = the 10 x 18 halo tile of 4 quads
static ptd_status tc_make_map_act(CUtensorMap* map, const DnTensor& t) {
    PFN_encodeTiled enc = tc_get_encode();
    if (!enc) PTD_FAIL(PTD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t epv = 16 / t.esize;                     // elements per 16-byte channel vector
    cuuint64_t dims[3] = {(cuuint64_t)t.W * epv, (cuuint64_t)t.rows + 2, (cuuint64_t)t.nvec()};
    cuuint64_t strides[2] = {(cuuint64_t)t.W * 16, (cuuint64_t)(t.rows + 2) * t.W * 16};
    cuuint32_t box[3] = {(cuuint32_t)(TC_HALO_W * epv), TC_HALO_H, 4};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, t.esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)t.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PTD_FAIL(PTD_ERR_CUDA, "cuTensorMapEncodeTiled(activation %dx%dx%d) failed: %d", t.rows, t.W, t.cp, (int)r);
    return PTD_OK;
}

inline void tc_plan_destroy(TcConvPlan& plan) { plan.valid = 0; }

// kernel variants: (operand type, output format) in {tf32 -> fp32, tf32 -> fp32 hi/lo, f16 -> fp32, f16 -> fp16, f16 -> fp16 hi/lo} x linked
typedef void (*TcKernel)(const TcParams);
static TcKernel tc_kernel_variant(int combo, int linked) {
    switch (combo * 2 + linked) {
        case 0: return conv_tc_kernel<false, false, 0>; case 1: return conv_tc_kernel<false, true, 0>;
        case 2: return conv_tc_kernel<false, false, 2>; case 3: return conv_tc_kernel<false, true, 2>;
        case 4: return conv_tc_kernel<true, false, 0>;  case 5: return conv_tc_kernel<true, true, 0>;
        case 6: return conv_tc_kernel<true, false, 1>;  case 7: return conv_tc_kernel<true, true, 1>;
        case 8: return conv_tc_kernel<true, false, 3>;  default: return conv_tc_kernel<true, true, 3>;
    }
}
static int tc_kernel_combo(bool half, int fmt) { return half ? (fmt == 0 ? 2 : (fmt == 1 ? 3 : 4)) : (fmt == 0 ? 0 : 1); }

// w9: [9][cinp][coutp] fp32 (padded, zero-filled); builds the packed tf32 weights, the tensor maps and the launch shape.
inline ptd_status tc_plan_create(const TcConvDesc& d, const std::vector<float>& w9, int cinp, TcConvPlan& plan, std::vector<void*>& allocs) {
    const int coutp = d.out.cp;
    const int c0p = d.src0.cp, c1p = d.src1.base ? d.src1.cp : 0;
    if (coutp % 16 || coutp < 16 || coutp > TC_ACC_COLS || c0p % 16 || c1p % 16 || c0p + c1p != cinp)
        PTD_FAIL(PTD_ERR_UNSUPPORTED, "tc conv: channel padding %d/%d -> %d unsupported", c0p, c1p, coutp);
    const bool half = d.src0.esize == 2;
    if (d.src1.base && d.src1.esize != d.src0.esize) PTD_FAIL(PTD_ERR_ARG, "tc conv: concat sources differ in element type");
    const int CH = half ? 32 : 16;                           // channels per chunk (4 channel vectors)
    TcParams& p = plan.p;
    memset(&p, 0, sizeof p);
    p.half = half ? 1 : 0;
    p.v0 = d.src0.nvec(); p.v1 = d.src1.base ? d.src1.nvec() : 0;
    p.n0 = (c0p + CH - 1) / CH; p.n1 = (c1p + CH - 1) / CH;
    const int nch = p.n0 + p.n1;
    // tile domain = the output's rows at source resolution; the sources hold at least rows [src_yoff, src_yoff + Hs)
    const int Hs = d.upsample ? d.out.rows / 2 : d.out.rows, Ws = d.src0.W;
    if ((d.upsample ? d.out.W != 2 * Ws : d.out.W != Ws) || d.src_yoff < 0 || d.src0.rows < Hs + d.src_yoff)
        PTD_FAIL(PTD_ERR_ARG, "tc conv: source %dx%d (+%d) does not cover output %dx%d", d.src0.rows, Ws, d.src_yoff, d.out.rows, d.out.W);
    if (d.src1.base && (d.src1.rows != d.src0.rows || d.src1.W != Ws)) PTD_FAIL(PTD_ERR_ARG, "tc conv: concat sources differ in size");
    p.Hs = Hs; p.Ws = Ws; p.out_mul = d.upsample ? 2 : 1; p.src_yoff = d.src_yoff;
    p.nphases = d.upsample ? 4 : 1; p.ntaps = d.upsample ? 4 : 9;
    p.coutp = coutp; p.scale = d.scale; p.shift = d.shift; p.bias = d.bias; p.lrelu_first = d.lrelu_first; p.round_out = d.round_out;
    p.out = d.out; p.pool_out = d.pool_out;
    p.tiles_x = (Ws + TC_TILE_W - 1) / TC_TILE_W; p.tiles_y = (Hs + TC_TILE_H - 1) / TC_TILE_H;
    p.total_items = p.tiles_x * p.tiles_y * p.nphases;
    p.x3 = d.src0.lo_off ? 1 : 0;
    if (p.x3 && d.src1.base && !d.src1.lo_off) PTD_FAIL(PTD_ERR_ARG, "tc conv: split-operand modes need hi/lo copies of both sources");
    if (nch > 32) PTD_FAIL(PTD_ERR_UNSUPPORTED, "tc conv: %d channel chunks (> 32)", nch);
    const int nblk = p.x3 ? 2 : 1;                           // weight blocks per (phase, chunk): hi [, lo]
    p.b_stage_bytes = (uint32_t)(p.ntaps * coutp * 64);
    p.w_total_bytes = (uint32_t)(p.nphases * nch * nblk) * p.b_stage_bytes;
    // shared memory plan: resident weights when they leave room for >= 4 A stages; the split modes fuse hi x hi | hi x lo into one MMA
    // (see mma_role) whenever a stage can carry a chunk's whole [hi | lo] slab and still leave three stages
    const size_t budget = TC_SMEM_BUDGET;
    p.b_load_bytes = p.b_stage_bytes;
    p.fused = 0;
    if ((size_t)p.w_total_bytes + 4 * TC_A_BYTES <= budget) {
        p.resident = 1;
        p.stages = (int)std::min<size_t>(TC_MAX_STAGES, (budget - p.w_total_bytes) / TC_A_BYTES);
        p.fused = p.x3;
    } else {
        p.resident = 0;
        if (p.x3 && 2 * coutp <= 256 && budget / (TC_A_BYTES + 2 * (size_t)p.b_stage_bytes) >= 3) { p.fused = 1; p.b_load_bytes = 2 * p.b_stage_bytes; }
        p.stages = (int)std::min<size_t>(TC_MAX_STAGES, budget / (TC_A_BYTES + p.b_load_bytes));
        if (p.stages < 2) PTD_FAIL(PTD_ERR_UNSUPPORTED, "tc conv: a stage of %u B does not fit twice in shared memory", TC_A_BYTES + p.b_load_bytes);
    }
    if (const char* e = getenv("PTD_DN_NO_FUSED_SPLIT")) { if (atoi(e) > 0 && p.fused) { p.fused = 0; if (!p.resident) { p.b_load_bytes = p.b_stage_bytes; p.stages = (int)std::min<size_t>(TC_MAX_STAGES, budget / (TC_A_BYTES + p.b_load_bytes)); } } }
    // accumulators (see TcParams): buffers of nseg main accumulators + (split modes) the correction accumulator(s), at least two buffers in 512 columns
    {
        const int max_accs = (TC_TMEM_COLS / 2) / coutp;
        if (p.fused) {                                        // pairs [main | correction]
            p.nseg = std::max(1, std::min(std::min(2, max_accs / 2), nch));
            p.buf_cols = 2 * p.nseg * coutp;
            p.corr_mask = 0xaaaaaaaau & ((1u << (2 * p.nseg)) - 1u);
        } else {
            p.nseg = p.x3 ? std::max(1, std::min(std::min(TC_MAX_SEGS, max_accs - 1), nch)) : 1;
            p.buf_cols = (p.nseg + p.x3) * coutp;
            p.corr_mask = p.x3 ? 1u << p.nseg : 0u;
        }
        p.nbuf = std::min(TC_MAX_BUFS, TC_TMEM_COLS / p.buf_cols);
        if (p.nbuf < 2) PTD_FAIL(PTD_ERR_UNSUPPORTED, "tc conv: %d columns of accumulators do not fit twice in tensor memory", p.buf_cols);
        p.seg_start = 0;
        for (int c = 1; c < nch; ++c)
            if ((c * p.nseg) / nch != ((c - 1) * p.nseg) / nch) p.seg_start |= 1u << c;
        p.corr_scale = half ? 1.0f / 2048.0f : 1.0f;
        if (const char* e = getenv("PTD_DN_DEBUG")) p.dbg = atoi(e);
    }
    // packed weights: [phase][chunk][tap][kstep j][k vector][n][16 bytes = 4 tf32 / 8 fp16]  (see make_desc_nosw); built as 16-bit
    // or 32-bit words in one byte buffer
    std::vector<float> pack((size_t)p.w_total_bytes / 4, 0.f);
    __half* pack_h = reinterpret_cast<__half*>(pack.data());
    for (int ph = 0; ph < p.nphases; ++ph) {
        for (int t = 0; t < p.ntaps; ++t) {
            int kys[3], kxs[3], nky = 0, nkx = 0;
            if (!d.upsample) {
                p.dy[ph][t] = t / 3 - 1; p.dx[ph][t] = t % 3 - 1;
                kys[nky++] = t / 3; kxs[nkx++] = t % 3;
            } else {
                // output (2y+a, 2x+b) of conv3x3(nearest_up2(S)): rows {y-1 | y} for a = 0 (ky = 0 | 1,2), {y | y+1} for a = 1 (ky = 0,1 | 2)
                const int a = ph >> 1, b = ph & 1, ti = t >> 1, tj = t & 1;
                p.dy[ph][t] = a == 0 ? ti - 1 : ti; p.dx[ph][t] = b == 0 ? tj - 1 : tj;
                if (a == 0) { if (ti == 0) kys[nky++] = 0; else { kys[nky++] = 1; kys[nky++] = 2; } }
                else { if (ti == 0) { kys[nky++] = 0; kys[nky++] = 1; } else kys[nky++] = 2; }
                if (b == 0) { if (tj == 0) kxs[nkx++] = 0; else { kxs[nkx++] = 1; kxs[nkx++] = 2; } }
                else { if (tj == 0) { kxs[nkx++] = 0; kxs[nkx++] = 1; } else kxs[nkx++] = 2; }
            }
            for (int c = 0; c < cinp; ++c)
                for (int n = 0; n < coutp; ++n) {
                    float s = 0.f;
                    for (int i = 0; i < nky; ++i)
                        for (int j = 0; j < nkx; ++j) s += w9[((size_t)(kys[i] * 3 + kxs[j]) * cinp + c) * coutp + n];
                    // position of padded-concat channel c: source, chunk of that source, K step, vector, element
                    const int cl = c < c0p ? c : c - c0p, chunk = (c < c0p ? 0 : p.n0) + cl / CH, r = cl % CH;
                    const int epv = CH / 4, kstep = r / (2 * epv), kq = (r % (2 * epv)) / epv, ke = r % epv;
                    // unfused: [phase][chunk][hi | lo block][tap][kstep][k vector][n]; fused split: [phase][chunk][tap][kstep][k vector][hi n .. | lo n ..]
                    const size_t slab = (size_t)((t * 2 + kstep) * 2 + kq);
                    const size_t off = p.fused ? (((size_t)(ph * nch + chunk) * p.ntaps * 4 + slab) * (size_t)(2 * coutp) + (size_t)n) * epv + ke
                                               : (((size_t)((ph * nch + chunk) * nblk) * p.ntaps * 4 + slab) * (size_t)coutp + (size_t)n) * epv + ke;
                    const size_t lo_off = p.fused ? (size_t)coutp * epv : (size_t)p.b_stage_bytes / (half ? 2 : 4);      // elements from a weight's hi to its lo copy
                    if (half) {
                        const __half hi = __float2half_rn(s);
                        pack_h[off] = hi;
                        if (p.x3) pack_h[off + lo_off] = __float2half_rn((s - __half2float(hi)) * 2048.0f);   // lo copy, scaled like the activations' lo copy
                    } else {
                        const float hi = host_round_tf32(s);
                        pack[off] = hi;
                        if (p.x3) pack[off + lo_off] = host_round_tf32(s - hi);
                    }
                }
        }
    }
    void* dw = (void*)d.shared_wpack;
    if (!dw) {
        if (cudaMalloc(&dw, pack.size() * 4) != cudaSuccess) PTD_FAIL(PTD_ERR_CUDA, "tc conv: cudaMalloc(weights) failed");
        allocs.push_back(dw);
        cudaMemcpy(dw, pack.data(), pack.size() * 4, cudaMemcpyHostToDevice);
    }
    plan.d_wpack = (float*)dw;
    p.wpack = plan.d_wpack;
    ptd_status rc = tc_make_map_act(&p.mapA0, d.src0);
    if (rc != PTD_OK) return rc;
    if (p.n1) { rc = tc_make_map_act(&p.mapA1, d.src1); if (rc != PTD_OK) return rc; }
    else p.mapA1 = p.mapA0;
    p.mapA0lo = p.mapA0; p.mapA1lo = p.mapA1;
    if (p.x3) {
        DnTensor t = d.src0; t.base += t.lo_off;
        rc = tc_make_map_act(&p.mapA0lo, t);
        if (rc != PTD_OK) return rc;
        if (p.n1) { t = d.src1; t.base += t.lo_off; rc = tc_make_map_act(&p.mapA1lo, t); if (rc != PTD_OK) return rc; }
    }
    // two MMA issuers need a stage ring each (>= 2 stages) and an even number of accumulator buffers (see mma_role)
    p.nissue = (p.stages >= 4 && p.nbuf >= 2) ? TC_MMA_WARPS : 1;
    if (const char* e = getenv("PTD_DN_ISSUERS")) { if (atoi(e) == 1) p.nissue = 1; }
    if (p.nissue == 2) { p.stages &= ~1; p.nbuf &= ~1; }
    const size_t b_bytes = p.resident ? (size_t)p.w_total_bytes : (size_t)p.stages * p.b_load_bytes;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    plan.grid = p.total_items < sms ? p.total_items : sms;
    plan.smem = (size_t)p.stages * TC_A_BYTES + ((b_bytes + 15) & ~(size_t)15) + 1024 + 512 + TC_ACC_COLS * 16;
    for (int v = 0; v < 10; ++v)
        if (cudaFuncSetAttribute(tc_kernel_variant(v >> 1, v & 1), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)) != cudaSuccess)
            PTD_FAIL(PTD_ERR_CUDA, "tc conv: cannot reserve shared memory: %s", cudaGetErrorString(cudaGetLastError()));
    plan.valid = 1;
    return PTD_OK;
}

inline ptd_status tc_conv_launch(TcConvPlan& plan, cudaStream_t st, int* launches, bool* pooled, int sm_limit = 0) {
    if (!plan.valid) PTD_FAIL(PTD_ERR_STATE, "tc conv: plan not built");
    const int fmt = tc::tensor_format(plan.p.out);
    if (plan.p.pool_out.base && tc::tensor_format(plan.p.pool_out) != fmt) PTD_FAIL(PTD_ERR_STATE, "tc conv: output and pooled output differ in storage format");
    if ((plan.p.half && fmt == 2) || (!plan.p.half && (fmt == 1 || fmt == 3))) PTD_FAIL(PTD_ERR_STATE, "tc conv: operand type / output format combination not built");
    TcKernel kern = tc_kernel_variant(tc_kernel_combo(plan.p.half != 0, fmt), plan.p.linked);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(sm_limit > 0 && sm_limit < plan.grid ? sm_limit : plan.grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = plan.smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = plan.p.pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, plan.p);
    if (e != cudaSuccess) PTD_FAIL(PTD_ERR_CUDA, "tc conv: cudaLaunchKernelEx failed: %s", cudaGetErrorString(e));
    if (launches) ++*launches;
    if (pooled) *pooled = plan.p.pool_out.base != nullptr;
    return PTD_OK;
}
