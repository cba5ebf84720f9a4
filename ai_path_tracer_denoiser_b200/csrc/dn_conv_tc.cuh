// Tensor-core conv engine of the denoiser: shifted-tap implicit GEMM on tcgen05 (sm_100a).
//
// One conv layer = GEMM  D[pixel, cout] = sum over (tap, cin) A[pixel + tap offset, cin] * B[cout, (tap, cin)]:
//   * M tile  = 128 output pixels = an 8 x 16 patch of the NHWC activation (pixel = TMEM lane),
//   * N       = padded Cout (16 .. 112, one tcgen05.mma N),
//   * K loop  = taps x 16-channel chunks; every (tap, chunk) is one pipeline stage:
//                 A: one TMA box {16 ch, 16 px, 8 rows} of the source tensor shifted by the tap offset - the image border
//                    (conv padding = 1) is TMA out-of-bounds zero fill, no halo copies, no im2col buffer;
//                 B: one TMA box {16 k, Cout} of the pre-packed weights;
//               both land K-major with the 64-byte swizzle, exactly the canonical UMMA layout, and feed two
//               tcgen05.mma.cta_group::1.kind::tf32 (K = 8 each) that accumulate in TMEM (fp32).
//   * concat (model.py:67,136-140) = the K loop walks two tensor maps; nearest upsample x2 (model.py:40) = four 2x2-tap
//     phase GEMMs on the half-resolution sources with the 3x3 weights pre-summed per phase (2.25x fewer MACs);
//   * epilogue (4 warps, thread = pixel): tcgen05.ld 32x32b.x16 -> bias/BatchNorm(eval)/LeakyReLU -> NHWC float4 stores,
//     MaxPool2d(2) fused through two warp shuffles (the 2x2 window lives in one warp), activations rounded to tf32 (RN)
//     so the next layer's operand truncation is exact.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (one lane), 2 = TMEM allocator, 4..7 = epilogue.  Pipelines: 8 smem stages
// (full/empty mbarriers) and 2 TMEM accumulator buffers (tmem_full/tmem_empty), persistent CTAs, one per SM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include "ptd_internal.h"

#define TC_STAGES 8
#define TC_STAGE_BYTES 16384            // A 8192 + B <= 7168, 1024-aligned
#define TC_A_BYTES 8192
#define TC_TILE_H 8
#define TC_TILE_W 16
#define TC_THREADS 256
#define TC_TMEM_COLS 256
#define TC_ACC_COLS 128

struct TcConvDesc {
    const float* src0; const float* src1; int c0p, c1p; int upsample; int H, W; int coutp; float* out; bool lrelu_first;
    const float* scale; const float* shift; const float* bias; float* pool_out;
    bool round_out = true;               // round stored activations to tf32 (all layers but the last)
};

struct __align__(64) TcParams {
    CUtensorMap mapA0, mapA1, mapB;
    int n0, n1;                          // 16-channel chunks of source 0 / source 1
    int ntaps, nphases;
    int dy[4][9], dx[4][9];
    int tiles_x, tiles_y, total_items;
    int Hs, Ws;                          // tile domain (= source resolution)
    int out_stride, Wout;
    int coutp;
    const float* scale; const float* shift; const float* bias;
    int lrelu_first, round_out;
    float* out; float* pool_out;
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
// K-major operand tile with 64-byte swizzle: rows 64 B apart, 8-row groups 512 B apart (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(512 >> 4) << 32;                     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version
    d |= (uint64_t)4 << 61;                              // layout type SWIZZLE_64B
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
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
}  // namespace tc

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = p.n0 + p.n1;
    const int nstages_item = p.ntaps * nchunks;
    const uint32_t stage_tx = TC_A_BYTES + (uint32_t)p.coutp * 64u;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&p.mapA0);
        if (p.n1) tc::prefetch_tmap(&p.mapA1);
        tc::prefetch_tmap(&p.mapB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer ==================================================================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int ph = item % p.nphases, tile = item / p.nphases;
                const int x0 = (tile % p.tiles_x) * TC_TILE_W, y0 = (tile / p.tiles_x) * TC_TILE_H;
                for (int t = 0; t < p.ntaps; ++t) {
                    const int yy = y0 + p.dy[ph][t], xx = x0 + p.dx[ph][t];
                    for (int c = 0; c < nchunks; ++c) {
                        tc::mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * TC_STAGE_BYTES;
                        tc::mbar_expect_tx(&full[stage], stage_tx);
                        if (c < p.n0) tc::tma_load_3d(sa, &p.mapA0, &full[stage], c * 16, xx, yy);
                        else tc::tma_load_3d(sa, &p.mapA1, &full[stage], (c - p.n0) * 16, xx, yy);
                        tc::tma_load_2d(sa + TC_A_BYTES, &p.mapB, &full[stage], 0, ((ph * p.ntaps + t) * nchunks + c) * p.coutp);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single lane) =========================================================================
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = tf32, both K-major, N = coutp, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.coutp >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * TC_ACC_COLS;
                for (int s = 0; s < nstages_item; ++s) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::fence_after_sync();
                    const uint32_t sa = tc::smem_u32(smem + stage * TC_STAGE_BYTES);
                    const uint64_t adesc = tc::make_desc_sw64(sa), bdesc = tc::make_desc_sw64(sa + TC_A_BYTES);
                    tc::mma_tf32(d_tmem, adesc, bdesc, idesc, s > 0 ? 1u : 0u);
                    tc::mma_tf32(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);            // +32 B along K inside the swizzle row
                    tc::mma_commit(&empty[stage]);                                    // frees the smem stage when the MMAs retire
                    if (s == nstages_item - 1) tc::mma_commit(&tmem_full[acc]);       // accumulator complete -> epilogue
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> bias/BN/LeakyReLU -> NHWC ============================================
        const int q = warp & 3;                                    // TMEM lane quarter this warp may read
        const int m = q * 32 + lane;                               // pixel of the tile == TMEM lane
        const int ty = m >> 4, tx = m & 15;
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const int ph = item % p.nphases, tile = item / p.nphases;
            const int x = (tile % p.tiles_x) * TC_TILE_W + tx, y = (tile / p.tiles_x) * TC_TILE_H + ty;
            const bool valid = x < p.Ws && y < p.Hs;
            const int oy = p.out_stride * y + (p.nphases > 1 ? (ph >> 1) : 0), ox = p.out_stride * x + (p.nphases > 1 ? (ph & 1) : 0);
            float* orow = p.out + ((size_t)oy * p.Wout + ox) * p.coutp;
            float* prow = p.pool_out ? p.pool_out + ((size_t)(y >> 1) * (p.Wout >> 1) + (x >> 1)) * p.coutp : nullptr;
            tc::mbar_wait(&tmem_full[acc], acc_phase);
            tc::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * TC_ACC_COLS;
            for (int c0 = 0; c0 < p.coutp; c0 += 16) {
                uint32_t r[16];
                tc::tmem_ld16(taddr + c0, r);
                tc::tmem_ld_wait();
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float a = __uint_as_float(r[j]);
                    const float sc = __ldg(&p.scale[c0 + j]), sh = __ldg(&p.shift[c0 + j]);
                    float v;
                    if (p.lrelu_first) { v = a + __ldg(&p.bias[c0 + j]); v = v > 0.f ? v : 0.1f * v; v = fmaf(v, sc, sh); }
                    else { v = fmaf(a, sc, sh); v = v > 0.f ? v : 0.1f * v; }
                    o[j] = p.round_out ? tc::round_tf32(v) : v;
                }
                if (valid) {
                    float4* d = reinterpret_cast<float4*>(orow + c0);
                    d[0] = make_float4(o[0], o[1], o[2], o[3]); d[1] = make_float4(o[4], o[5], o[6], o[7]);
                    d[2] = make_float4(o[8], o[9], o[10], o[11]); d[3] = make_float4(o[12], o[13], o[14], o[15]);
                }
                if (p.pool_out) {                                  // MaxPool2d(2): partners are lanes ^1 (x) and ^16 (y)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float v = fmaxf(o[j], __shfl_xor_sync(0xffffffffu, o[j], 1));
                        o[j] = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
                    }
                    if (valid && !(tx & 1) && !(ty & 1)) {
                        float4* d = reinterpret_cast<float4*>(prow + c0);
                        d[0] = make_float4(o[0], o[1], o[2], o[3]); d[1] = make_float4(o[4], o[5], o[6], o[7]);
                        d[2] = make_float4(o[8], o[9], o[10], o[11]); d[3] = make_float4(o[12], o[13], o[14], o[15]);
                    }
                }
            }
            tc::fence_before_sync();
            tc::mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) {
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

static ptd_status tc_make_map_act(CUtensorMap* map, const float* base, int cp, int W, int H) {
    PFN_encodeTiled enc = tc_get_encode();
    if (!enc) PTD_FAIL(PTD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)cp, (cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[2] = {(cuuint64_t)cp * 4, (cuuint64_t)W * cp * 4};
    cuuint32_t box[3] = {16, TC_TILE_W, TC_TILE_H};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PTD_FAIL(PTD_ERR_CUDA, "cuTensorMapEncodeTiled(activation %dx%dx%d) failed: %d", H, W, cp, (int)r);
    return PTD_OK;
}

inline void tc_plan_destroy(TcConvPlan& plan) { plan.valid = 0; }

// w9: [9][cinp][coutp] fp32 (padded, zero-filled); builds the packed tf32 weights, the tensor maps and the launch shape.
inline ptd_status tc_plan_create(const TcConvDesc& d, const std::vector<float>& w9, int cinp, TcConvPlan& plan, std::vector<void*>& allocs) {
    if (d.coutp % 16 || d.coutp < 16 || d.coutp > TC_ACC_COLS || d.c0p % 16 || d.c1p % 16)
        PTD_FAIL(PTD_ERR_UNSUPPORTED, "tc conv: channel padding %d/%d -> %d unsupported", d.c0p, d.c1p, d.coutp);
    TcParams& p = plan.p;
    memset(&p, 0, sizeof p);
    p.n0 = d.c0p / 16; p.n1 = d.c1p / 16;
    const int nch = p.n0 + p.n1;
    const int Hs = d.upsample ? d.H / 2 : d.H, Ws = d.upsample ? d.W / 2 : d.W;
    p.Hs = Hs; p.Ws = Ws; p.Wout = d.W; p.out_stride = d.upsample ? 2 : 1;
    p.nphases = d.upsample ? 4 : 1; p.ntaps = d.upsample ? 4 : 9;
    p.coutp = d.coutp; p.scale = d.scale; p.shift = d.shift; p.bias = d.bias; p.lrelu_first = d.lrelu_first; p.round_out = d.round_out;
    p.out = d.out; p.pool_out = d.pool_out;
    p.tiles_x = (Ws + TC_TILE_W - 1) / TC_TILE_W; p.tiles_y = (Hs + TC_TILE_H - 1) / TC_TILE_H;
    p.total_items = p.tiles_x * p.tiles_y * p.nphases;
    // taps and packed weights  B[((phase * ntaps + tap) * nch + chunk) * coutp + n][16]
    std::vector<float> pack((size_t)p.nphases * p.ntaps * nch * d.coutp * 16, 0.f);
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
                for (int n = 0; n < d.coutp; ++n) {
                    float s = 0.f;
                    for (int i = 0; i < nky; ++i)
                        for (int j = 0; j < nkx; ++j) s += w9[((size_t)(kys[i] * 3 + kxs[j]) * cinp + c) * d.coutp + n];
                    pack[((((size_t)ph * p.ntaps + t) * nch + c / 16) * d.coutp + n) * 16 + c % 16] = host_round_tf32(s);
                }
        }
    }
    void* dw = nullptr;
    if (cudaMalloc(&dw, pack.size() * 4) != cudaSuccess) PTD_FAIL(PTD_ERR_CUDA, "tc conv: cudaMalloc(weights) failed");
    allocs.push_back(dw);
    cudaMemcpy(dw, pack.data(), pack.size() * 4, cudaMemcpyHostToDevice);
    plan.d_wpack = (float*)dw;
    ptd_status rc = tc_make_map_act(&p.mapA0, d.src0, d.c0p, Ws, Hs);
    if (rc != PTD_OK) return rc;
    if (p.n1) { rc = tc_make_map_act(&p.mapA1, d.src1, d.c1p, Ws, Hs); if (rc != PTD_OK) return rc; }
    else p.mapA1 = p.mapA0;
    {
        PFN_encodeTiled enc = tc_get_encode();
        cuuint64_t dims[2] = {16, (cuuint64_t)p.nphases * p.ntaps * nch * d.coutp};
        cuuint64_t strides[1] = {64};
        cuuint32_t box[2] = {16, (cuuint32_t)d.coutp};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dw, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) PTD_FAIL(PTD_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    plan.grid = p.total_items < sms ? p.total_items : sms;
    plan.smem = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 + 256;
    if (cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) != cudaSuccess)
        PTD_FAIL(PTD_ERR_CUDA, "tc conv: cannot reserve %zu B of shared memory: %s", plan.smem, cudaGetErrorString(cudaGetLastError()));
    plan.valid = 1;
    return PTD_OK;
}

inline ptd_status tc_conv_launch(TcConvPlan& plan, cudaStream_t st, int* launches, bool* pooled) {
    if (!plan.valid) PTD_FAIL(PTD_ERR_STATE, "tc conv: plan not built");
    conv_tc_kernel<<<plan.grid, TC_THREADS, plan.smem, st>>>(plan.p);
    if (launches) ++*launches;
    if (pooled) *pooled = plan.p.pool_out != nullptr;
    return PTD_OK;
}
