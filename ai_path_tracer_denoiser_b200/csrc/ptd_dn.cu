// HP-2: forward pass of the recurrent denoising autoencoder, B200-native.
//
// Replaces network_prediction_faster_version() / torch::jit Module.forward (Inference/src/main.cpp:101-118) for the
// network defined in training/recurrent_autoencoder_model.py:8-142 - 28 x (conv3x3 pad 1 + BatchNorm(eval) + LeakyReLU 0.1),
// 5 x MaxPool2, 5 x nearest Upsample x2, 11 x channel concat, 6 recurrent hidden states - with eval-mode BN and carried
// hidden state (SURVEY.md decisions D1, D3).
//
// Data layout in HBM ("CHW4"): every activation is fp32 channel-quad planes [C/4][rows + 2][W][4] with the channel count padded
// to a multiple of 16 (10->16, 32, 43->48, 57->64, 76->80, 101->112, 3->16; pad channels are kept at exactly 0) and one zero
// apron row above and below the image (conv padding in y; the halo slots of the multi-GPU row-strip mode).  A pixel's four
// channels of a quad are one 16-byte vector and a row of a quad is contiguous, so (a) CUDA-core loads / stores are float4 and
// coalesced along x and (b) a TMA box of a halo tile lands in shared memory directly in the canonical no-swizzle K-major UMMA
// layout (dn_conv_tc.cuh).  Concats and the decoder's upsample are never materialised (the convs read two sources / index at
// half resolution); conv bias, BN and LeakyReLU are folded into the conv epilogue.  Two conv engines share the buffers and
// the layer graph:
//   PTD_DN_FP32  conv3x3_fp32   - implicit-GEMM on CUDA cores (FFMA), strict-parity path;
//   PTD_DN_TF32  dn_conv_tc.cuh - TMA-staged tiles + tcgen05.mma kind::tf32 with TMEM accumulators.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <unistd.h>
#include <algorithm>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include "ptd_internal.h"
#include "dn_layers.h"
#include "dn_conv_tc.cuh"

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ptd_set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return PTD_ERR_CUDA; } } while (0)

// ---- CUDA-core conv engine --------------------------------------------------------------------------------------------
// Block = 128 threads -> 8 x 16 output pixels x 32 output channels; thread = 4 vertically adjacent pixels x 8 channels.
// K loop over chunks of 8 input channels: the 10 x 18 halo tile of the chunk and the 9 x 8 x 32 weights are staged in shared
// memory; per (channel, kx) a thread loads 6 inputs + 3 x 8 weights for 96 FMAs.
#define FC_TH 8
#define FC_TW 16
#define FC_CK 8
#define FC_BN 32
#define FC_RS 20          // smem row stride (18 used): keeps the two half-warps on disjoint banks
#define FC_PS (10 * FC_RS)

struct FpConvArgs {
    DnTensor src0, src1;                    // CHW4 (src1.base may be null)
    int upsample;                           // sources live at (H/2, W/2): nearest x2 (model.py:40)
    int H, W;                               // output (= conv input) resolution
    const float* w;                         // [9][c0p + c1p][coutp]
    const float* scale; const float* shift; const float* bias;   // [coutp]
    int order;                              // 0: lrelu(scale*acc + shift)   1: scale*lrelu(acc + bias) + shift
    DnTensor out;
};

// RAW = 1 (PTD_DN_FP32_BATCH_STATS): store the value BatchNorm will see - acc + bias (order 2) or lrelu(acc + bias) (order 3) - and leave
// the normalisation to bn_batch_stats / bn_batch_apply, which need the whole layer's output first.
template <int RAW>
__global__ void __launch_bounds__(128) conv3x3_fp32(const FpConvArgs a) {
    __shared__ float s_in[FC_CK * FC_PS];
    __shared__ __align__(16) float s_w[9 * FC_CK * FC_BN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * FC_TW, y0 = blockIdx.y * FC_TH, co0 = blockIdx.z * FC_BN;
    const int col = lane & 15, rbase = (lane >> 4) * 4;     // thread's pixels: rows rbase..rbase+3 of the tile, column col
    const int cg = warp * 8;                                  // thread's 8 output channels inside the 32-wide tile
    const int c0p = a.src0.cp, c1p = a.src1.base ? a.src1.cp : 0, coutp = a.out.cp;
    const int cin = c0p + c1p;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < cin; c0 += FC_CK) {
        const DnTensor& src = c0 < c0p ? a.src0 : a.src1;
        const int cs = c0 < c0p ? c0 : c0 - c0p;
        __syncthreads();
        // halo tile: 10 x 18 pixels x 8 channels (two quads per pixel), zero outside the image (padding = 1)
        for (int i = tid; i < 10 * 18 * 2; i += 128) {
            const int half = i / (10 * 18), pix = i % (10 * 18), r = pix / 18, c = pix % 18;
            const int gy = y0 + r - 1, gx = x0 + c - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
                const int sy = a.upsample ? gy >> 1 : gy, sx = a.upsample ? gx >> 1 : gx;
                v = __ldg(reinterpret_cast<const float4*>(src.base + (size_t)(cs / 4 + half) * src.quad_stride() + ((size_t)(sy + 1) * src.W + sx) * 4));
            }
            float* d = s_in + (half * 4) * FC_PS + r * FC_RS + c;
            d[0] = v.x; d[FC_PS] = v.y; d[2 * FC_PS] = v.z; d[3 * FC_PS] = v.w;
        }
        // weights of this chunk: [9][8][32]
        for (int i = tid; i < 9 * FC_CK * FC_BN / 4; i += 128) {
            const int j4 = i % (FC_BN / 4), c = (i / (FC_BN / 4)) % FC_CK, tap = i / (FC_BN / 4 * FC_CK);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (co0 + j4 * 4 < coutp) v = __ldg(reinterpret_cast<const float4*>(a.w + ((size_t)tap * cin + c0 + c) * coutp + co0 + j4 * 4));
            reinterpret_cast<float4*>(s_w)[i] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int c = 0; c < FC_CK; ++c) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float in[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) in[r] = s_in[c * FC_PS + (rbase + r) * FC_RS + col + kx];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&s_w[((ky * 3 + kx) * FC_CK + c) * FC_BN + cg]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&s_w[((ky * 3 + kx) * FC_CK + c) * FC_BN + cg + 4]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float v = in[i + ky];
                        acc[i][0] = fmaf(v, w0.x, acc[i][0]); acc[i][1] = fmaf(v, w0.y, acc[i][1]);
                        acc[i][2] = fmaf(v, w0.z, acc[i][2]); acc[i][3] = fmaf(v, w0.w, acc[i][3]);
                        acc[i][4] = fmaf(v, w1.x, acc[i][4]); acc[i][5] = fmaf(v, w1.y, acc[i][5]);
                        acc[i][6] = fmaf(v, w1.z, acc[i][6]); acc[i][7] = fmaf(v, w1.w, acc[i][7]);
                    }
                }
            }
        }
    }
    const int co = co0 + cg;
    if (co >= coutp) return;
    float sc[8], sh[8], bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = a.scale[co + j]; sh[j] = a.shift[co + j]; bi[j] = a.bias[co + j]; }
    const int gx = x0 + col;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gy = y0 + rbase + i;
        if (gy >= a.H || gx >= a.W) continue;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (RAW) { float v = acc[i][j] + bi[j]; o[j] = (a.order == 3 && !(v > 0.f)) ? 0.1f * v : v; }
            else if (a.order == 0) { float v = fmaf(acc[i][j], sc[j], sh[j]); o[j] = v > 0.f ? v : 0.1f * v; }
            else { float v = acc[i][j] + bi[j]; v = v > 0.f ? v : 0.1f * v; o[j] = fmaf(v, sc[j], sh[j]); }
        }
        float* dst = a.out.base + (size_t)(co / 4) * a.out.quad_stride() + ((size_t)(gy + 1) * a.out.W + gx) * 4;
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dst + a.out.quad_stride()) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// MaxPool2d(2) on CHW4 (model.py:19): one thread = one output pixel of one channel quad
__global__ void maxpool2_chw4(const DnTensor in, const DnTensor out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nq = out.cp / 4, Ho = out.rows, Wo = out.W;
    if (i >= (size_t)nq * Ho * Wo) return;
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho), q = (int)(i / ((size_t)Wo * Ho));
    const float4* p = reinterpret_cast<const float4*>(in.base + (size_t)q * in.quad_stride()) + (size_t)(2 * y + 1) * in.W + 2 * x;
    const float4 a = p[0], b = p[1], d = p[in.W], e = p[in.W + 1];
    float4* o = reinterpret_cast<float4*>(out.base + (size_t)q * out.quad_stride()) + (size_t)(y + 1) * Wo + x;
    *o = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                     fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
}
// ---- PTD_DN_FP32_BATCH_STATS: BatchNorm with the statistics of the CURRENT activations (what the reference's TorchScript export
// actually runs: convert_to_torchscript.py:26-30 traces the model without .eval(), so nn.BatchNorm2d normalises with the batch
// mean and the biased batch variance of its input, N = 1 -> over the H x W pixels of the padded frame) ----
// stats[2 * c] = sum, stats[2 * c + 1] = sum of squares of channel c over the image rows (aprons excluded), in double
__global__ void __launch_bounds__(256) bn_batch_stats(const DnTensor t, double* __restrict__ stats) {
    __shared__ double s_red[8][8];
    const int q = blockIdx.y;
    const float4* plane = reinterpret_cast<const float4*>(t.base + (size_t)q * t.quad_stride()) + t.W;      // first image row
    const size_t n = (size_t)t.rows * t.W;
    double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = plane[i];
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        ss[0] += (double)v.x * v.x; ss[1] += (double)v.y * v.y; ss[2] += (double)v.z * v.z; ss[3] += (double)v.w * v.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s[k] += __shfl_xor_sync(0xffffffffu, s[k], o); ss[k] += __shfl_xor_sync(0xffffffffu, ss[k], o); }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { for (int k = 0; k < 4; ++k) { s_red[warp][k] = s[k]; s_red[warp][4 + k] = ss[k]; } }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += s_red[w][threadIdx.x];
        const int c = q * 4 + (threadIdx.x & 3);
        atomicAdd(&stats[2 * c + (threadIdx.x >> 2)], v);
    }
}
// in place: y = (x - mean) / sqrt(var + eps) * gamma + beta, then LeakyReLU when the layer is conv -> BN -> LReLU (model.py:24-26);
// the conv -> LReLU -> BN layer (model.py:30-32) got its LeakyReLU in the conv epilogue.  Pad channels have gamma = beta = 0.
__global__ void __launch_bounds__(256) bn_batch_apply(const DnTensor t, const double* __restrict__ stats, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, int lrelu_after) {
    const int q = blockIdx.y;
    float4* plane = reinterpret_cast<float4*>(t.base + (size_t)q * t.quad_stride()) + t.W;
    const size_t n = (size_t)t.rows * t.W;
    float mul[4], add[4], mu[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = q * 4 + k;
        const double mean = stats[2 * c] / (double)n, var = fmax(stats[2 * c + 1] / (double)n - mean * mean, 0.0);
        const float invstd = (float)(1.0 / sqrt(var + 1e-5));
        mu[k] = (float)mean; mul[k] = invstd * gamma[c]; add[k] = beta[c];
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = plane[i];
        v.x = fmaf(v.x - mu[0], mul[0], add[0]); v.y = fmaf(v.y - mu[1], mul[1], add[1]); v.z = fmaf(v.z - mu[2], mul[2], add[2]); v.w = fmaf(v.w - mu[3], mul[3], add[3]);
        if (lrelu_after) { v.x = v.x > 0.f ? v.x : 0.1f * v.x; v.y = v.y > 0.f ? v.y : 0.1f * v.y; v.z = v.z > 0.f ? v.z : 0.1f * v.z; v.w = v.w > 0.f ? v.w : 0.1f * v.w; }
        plane[i] = v;
    }
}

// planar G-buffer [10][H][W] -> CHW4 with 16 channels (4 quads): the strip's rows (frame rows row0 .. row0 + rows - 1), zero in the
// bottom / right padding (decision D3); thread = pixel.  Apron rows: zero at the frame border; where a neighbour strip exists
// it pushes its boundary row into OUR apron (and we push ours into its), then the last block raises the neighbours' flags -
// the same protocol as the conv epilogues (dn_conv_tc.cuh), so only the strip's own G-buffer rows need to be valid here.
struct PackLink { DnTensor up, down; uint32_t* sig_up; uint32_t* sig_down; uint32_t* done; uint32_t epoch; };
__global__ void pack_gbuffer(const float* __restrict__ g, int H, int W, int row0, const DnTensor out, int round_tf32, const PackLink link) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int Wp = out.W;
    if (i < (size_t)(out.rows + 2) * Wp) {
        const int x = (int)(i % Wp), br = (int)(i / Wp), y = row0 + br - 1;
        const bool apron = br == 0 || br == out.rows + 1;
        if (!(apron && ((br == 0 && link.up.base) || (br != 0 && link.down.base)))) {   // a neighbour owns that apron row
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = 0.f;
            if (!apron && x < W && y >= 0 && y < H) {
#pragma unroll
                for (int c = 0; c < 10; ++c) v[c] = g[(size_t)c * H * W + (size_t)y * W + x];
                if (round_tf32) {
#pragma unroll
                    for (int c = 0; c < 10; ++c) v[c] = tc::round_tf32(v[c]);
                }
            }
            if (out.esize == 2) {
#pragma unroll
                for (int c = 0; c < 10; ++c) v[c] = fminf(fmaxf(v[c], -65504.f), 65504.f);       // fp16 range (the depth plane is unbounded)
            }
            tc::store16(out, out.base + ((size_t)br * Wp + x) * 4, 0, v);
            if (br == 1 && link.up.base) tc::store16(link.up, link.up.base + ((size_t)(link.up.rows + 1) * Wp + x) * 4, 0, v);
            if (br == out.rows && link.down.base) tc::store16(link.down, link.down.base + (size_t)x * 4, 0, v);
        }
    }
    if (link.done) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t old = atomicAdd(link.done, 1u);
            if (old + 1u == gridDim.x) {
                *link.done = 0u;
                __threadfence_system();
                if (link.sig_up) tc::st_release_sys(link.sig_up, link.epoch);
                if (link.sig_down) tc::st_release_sys(link.sig_down, link.epoch);
            }
        }
    }
}
// CHW4 quad 0 -> planar [3][H][W]: frame rows [r0, r0 + nrows) of this strip (tensor row 1 == frame row r0), cropped to W
__global__ void unpack_rgb(const DnTensor in, int H, int W, int r0, int nrows, float* __restrict__ rgb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nrows * W) return;
    const int x = (int)(i % W), ry = (int)(i / W);
    const float4 v = *reinterpret_cast<const float4*>(in.base + ((size_t)(ry + 1) * in.W + x) * 4);
    const size_t o = (size_t)(r0 + ry) * W + x;
    rgb[o] = v.x; rgb[(size_t)H * W + o] = v.y; rgb[(size_t)2 * H * W + o] = v.z;
}
// CHW4 -> NCHW (hidden-state parity tap)
__global__ void chw4_to_nchw(const DnTensor in, int C, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int H = in.rows, W = in.W;
    if (i >= (size_t)C * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)(i / ((size_t)W * H));
    if (in.esize == 4) {
        const size_t o = (size_t)(c >> 2) * in.quad_stride() + ((size_t)(y + 1) * W + x) * 4 + (c & 3);
        out[i] = in.lo_off ? in.base[o] + in.base[o + in.lo_off] : in.base[o];
    } else {
        const float* v8 = in.base + (size_t)(c >> 3) * in.quad_stride() + ((size_t)(y + 1) * W + x) * 4;
        const float hi = __half2float(reinterpret_cast<const __half*>(v8)[c & 7]);
        out[i] = in.lo_off ? hi + __half2float(reinterpret_cast<const __half*>(v8 + in.lo_off)[c & 7]) * (1.0f / 2048.0f) : hi;      // 2xF16: lo is stored x 2^11
    }
}

// ---- weights ---------------------------------------------------------------------------------------------------------
static ptd_status read_ptdw(const char* path, std::map<std::string, std::vector<float>>& sd) {
    FILE* f = fopen(path, "rb");
    if (!f) PTD_FAIL(PTD_ERR_IO, "cannot open weight file '%s'", path);
    auto fail = [&](const char* why) { fclose(f); ptd_set_error("weight file '%s': %s", path, why); return PTD_ERR_PARSE; };
    char magic[4]; uint32_t ver = 0, n = 0;
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "PTDW", 4) != 0) return fail("bad magic (expected PTDW)");
    if (fread(&ver, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1 || ver != 1) return fail("unsupported version");
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t kl = 0, nd = 0;
        if (fread(&kl, 4, 1, f) != 1 || kl > 4096) return fail("truncated (name length)");
        std::string k(kl, '\0');
        if (fread(&k[0], 1, kl, f) != kl || fread(&nd, 4, 1, f) != 1 || nd > 8) return fail("truncated (name / ndim)");
        size_t cnt = 1;
        for (uint32_t d = 0; d < nd; ++d) { uint32_t s = 0; if (fread(&s, 4, 1, f) != 1) return fail("truncated (dims)"); cnt *= s; }
        if (cnt > (1u << 28)) return fail("tensor too large");
        std::vector<float> v(cnt);
        if (cnt && fread(v.data(), 4, cnt, f) != cnt) return fail("truncated (data)");
        sd[k] = std::move(v);
    }
    fclose(f);
    return PTD_OK;
}

// ---- handle ----------------------------------------------------------------------------------------------------------
// Row strips (multi-GPU): a handle owns padded rows [row0, row0 + rows) of the frame, rows % 32 == 0, i.e.
// rows >> l rows of every level-l tensor.  Every 3x3 conv needs one row of its inputs from the strip above and below; those
// rows are the apron rows of the CHW4 tensors.  The producing conv's epilogue stores its first / last output row straight
// into the neighbour GPU's apron (peer pointers from cudaIpcOpenMemHandle, NVLink) and the launch's last CTA raises a
// per-tensor flag there (st.release.sys); the consuming conv's TMA producer spins on its local flags (ld.acquire.sys) before
// the first load.  Flags carry the frame sequence number, so nothing is ever reset.  The six hidden states are double
// buffered (frame k reads parity k & 1, writes the other): a neighbour that runs ahead can then never overwrite an apron
// row that is still being read.  One strip covering the whole frame is the single-GPU case: aprons stay zero.
// Replicated levels: tiling stops at level DN_REPL_LEVEL (1/8 resolution).  Below it a strip would be a handful of rows and every
// layer would pay a cross-GPU wait for almost no work, so levels >= DN_REPL_LEVEL (encoder 4-5, bottleneck, decoder 5-4: 13 of the
// 28 convs, 3 % of the FLOPs) are computed IN FULL by every strip: encoder 3's fused max-pool stores its rows of the level-3 input
// into every strip's full-height tensor ("gather": peer stores to all strips + one flag per source strip), the 13 convs then run
// without any exchange, and decoder 3 reads its rows of the replicated level-3 tensors back into the strip (src_yoff).
// Default for strip handles; PTD_DN_REPL_LEVEL=6 at ptd_dn_create_strip time (same value on every rank) tiles every level instead.
// Both are bit-identical to the untiled run (tests/test_gpu_dn.py, tests/test_gpu_pt.py; multi-process: tools/check_frame_strips.py).
#define DN_MAX_TENSORS 48
#define DN_MAX_RANKS 8
#define DN_REPL_LEVEL 3                       /* the level replication starts at when enabled */
struct ptd_strip_info {                      // POD, exchanged between the ranks as bytes (ptd_dn_strip_export / _connect)
    unsigned char ipc[64];                   // cudaIpcMemHandle_t of the activation arena
    unsigned long long arena;                // the arena's address in the owner's process (same-process connections)
    int pid_tag, device, row0, rows, Hp, Wp, ntensors, reserved;
    unsigned long long tensor_off[DN_MAX_TENSORS];   // byte offset of tensor i in the arena
    int tensor_rows[DN_MAX_TENSORS];
    unsigned long long flags_off;            // uint32 flags[ntensors][2] (from up, from down), 32-byte stride
    unsigned long long gflags_off;           // uint32 gather flags[DN_MAX_RANKS] (rows of strip r have arrived), 32-byte stride
};

struct DnLayer {
    DnLayerSpec spec;
    int H, W;                       // output resolution (this strip)
    int src0, src1, out, pool;      // tensor ids (-1 = none); hidden-state ids are resolved per parity
    float* d_w9 = nullptr;          // [9][cin_p][coutp]  (CUDA-core engine)
    float* d_scale = nullptr; float* d_shift = nullptr; float* d_bias = nullptr;
    float* d_gamma = nullptr; float* d_beta = nullptr;       // PTD_DN_FP32_BATCH_STATS: the BatchNorm affine parameters, unfolded
    TcConvPlan tc[2];               // tensor-core engine plans, one per hidden-state parity
    uint32_t* d_done = nullptr;
};

struct ptd_dn {
    int device = 0; unsigned flags = 0;
    int H = 0, W = 0, Hp = 0, Wp = 0;           // frame, padded frame
    int row0 = 0, rows = 0;                     // this strip (padded rows)
    bool strip = false;
    std::vector<DnLayer> layers;
    std::vector<void*> allocs;
    unsigned char* arena = nullptr; size_t arena_bytes = 0;
    std::vector<DnTensor> tensors;              // id -> tensor (base inside the arena)
    std::vector<size_t> tensor_off;
    size_t flags_off = 0;
    int t_in16 = -1, t_hidden[6][2], t_final = -1;
    int hidden_c[6] = {0};
    float* d_gbuf = nullptr; float* d_rgb = nullptr;   // staging for the host-pointer entry point
    double* d_bn_stats = nullptr;                      // PTD_DN_FP32_BATCH_STATS: [2 * 128] per-channel sum / sum of squares of the layer in flight
    struct Pool { int in[2], out; };
    std::vector<Pool> pools;        // pools[k] follows layer 3k+2 (CUDA-core engine only; the TC engine pools in its epilogue)
    // the other strips of the frame, by rank (top strip = 0); up = rank - 1, down = rank + 1
    int rank = 0, nranks = 1;
    unsigned char* peer_arena[DN_MAX_RANKS] = {nullptr};
    bool peer_ipc[DN_MAX_RANKS] = {false};
    ptd_strip_info peer_info[DN_MAX_RANKS];
    bool has_peer[2] = {false, false};                     // a strip above / below exists
    size_t gflags_off = 0;
    bool pdl = false;                                      // PTD_DN_PDL=1: convs launched with programmatic stream serialization
    int repl_level = 6;                                    // levels >= this are replicated on every strip (6 = none; strips default to DN_REPL_LEVEL)
    int t_gather = -1;                                     // the gathered tensor (pooled output of encoder repl_level, full height)
    std::vector<int> tensor_level; std::vector<char> tensor_full;
    uint32_t epoch = 0;
    uint32_t* d_pack_done = nullptr;
    int parity = 0;
    int sm_limit = 0;                                      // > 0: at most this many conv CTAs per launch (the denoiser's SM partition, ptd_frame_submit)
    int inflight = 0;                                      // frames enqueued by ptd_frame_submit and not yet taken by ptd_frame_wait (they own this handle's state)
    int launches = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> events;          // events[i], events[i+1] bracket launch i
    std::vector<std::string> launch_names;
    int timed_launches = 0;
    uint32_t* flag(int tensor, int dir) const { return (uint32_t*)(arena + flags_off + ((size_t)tensor * 2 + dir) * 32); }
};

extern "C" void ptd_dn_destroy(ptd_dn* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int d = 0; d < DN_MAX_RANKS; ++d) if (h->peer_arena[d] && h->peer_ipc[d]) cudaIpcCloseMemHandle(h->peer_arena[d]);
    for (auto& L : h->layers) { tc_plan_destroy(L.tc[0]); tc_plan_destroy(L.tc[1]); }
    for (void* p : h->allocs) cudaFree(p);
    cudaFree(h->arena);
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    delete h;
}

static inline int cpad(int c) { return (c + 15) & ~15; }
static inline bool dn_cuda_core_engine(unsigned flags) { return flags == PTD_DN_FP32 || flags == PTD_DN_FP32_BATCH_STATS; }

static ptd_status dn_create(const char* weights_path, int H, int W, int row0, int rows, bool strip, int device, unsigned flags, ptd_dn** out) {
    if (!weights_path || !out || H <= 0 || W <= 0) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_create: bad argument");
    *out = nullptr;
    if (ptd_device_count() <= device || device < 0) PTD_FAIL(PTD_ERR_CUDA, "ptd_dn_create: CUDA device %d not available (no CPU fallback exists)", device);
    if (flags != PTD_DN_FP32 && flags != PTD_DN_TF32 && flags != PTD_DN_F16 && flags != PTD_DN_3XTF32 && flags != PTD_DN_2XF16 && flags != PTD_DN_FP32_BATCH_STATS) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_create: unknown flags %u", flags);
    const int Hp = (H + 31) / 32 * 32, Wp = (W + 31) / 32 * 32;
    if (!strip) { row0 = 0; rows = Hp; }
    if (row0 < 0 || rows <= 0 || row0 % 32 || rows % 32 || row0 + rows > Hp) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_create_strip: rows [%d, %d) must be multiples of 32 inside the padded frame of %d rows", row0, row0 + rows, Hp);
    if (strip && dn_cuda_core_engine(flags)) PTD_FAIL(PTD_ERR_UNSUPPORTED, "ptd_dn_create_strip: row strips need a tensor-core engine (PTD_DN_TF32 / PTD_DN_F16)");
    std::map<std::string, std::vector<float>> sd;
    ptd_status rc = read_ptdw(weights_path, sd);
    if (rc != PTD_OK) return rc;
    CUDA_TRY(cudaSetDevice(device));
    ptd_dn* h = new ptd_dn();
    h->device = device; h->flags = flags; h->H = H; h->W = W; h->Hp = Hp; h->Wp = Wp; h->row0 = row0; h->rows = rows; h->strip = strip;
    if (const char* e = getenv("PTD_DN_PDL")) h->pdl = atoi(e) > 0;
    if (strip) h->repl_level = DN_REPL_LEVEL;             // PTD_DN_REPL_LEVEL=6 (same on every rank): tile every level instead
    if (const char* e = getenv("PTD_DN_REPL_LEVEL")) { const int v = atoi(e); if (v >= 3 && v <= 6) h->repl_level = v; }
    auto fail = [&](ptd_status code) { ptd_dn_destroy(h); return code; };
    auto dalloc = [&](size_t floats) -> float* {
        void* p = nullptr;
        if (cudaMalloc(&p, floats * 4 + 256) != cudaSuccess) { ptd_set_error("ptd_dn_create: cudaMalloc(%zu B) failed: %s", floats * 4, cudaGetErrorString(cudaGetLastError())); return nullptr; }
        cudaMemset(p, 0, floats * 4 + 256);
        h->allocs.push_back(p);
        return (float*)p;
    };
#define DALLOC(var, floats) do { (var) = dalloc(floats); if (!(var)) return fail(PTD_ERR_CUDA); } while (0)
    DALLOC(h->d_gbuf, (size_t)10 * H * W);
    DALLOC(h->d_rgb, (size_t)3 * H * W);
    { float* dd = nullptr; DALLOC(dd, 64); h->d_pack_done = (uint32_t*)dd; }
    if (flags == PTD_DN_FP32_BATCH_STATS) { float* dd = nullptr; DALLOC(dd, 2 * 128 * 2); h->d_bn_stats = (double*)dd; }

    // ---- activation arena: every tensor the convs read or write, one allocation (one IPC handle per strip) ----
    size_t arena_bytes = 0;
    const int act_esize = (flags == PTD_DN_F16 || flags == PTD_DN_2XF16) ? 2 : 4;       // fp16 activations everywhere but the network's output
    auto tnew = [&](int c, int lvl, int esize = 0) -> int {
        DnTensor t;
        const bool full = strip && lvl >= h->repl_level;                // replicated level: every strip holds the whole tensor
        t.cp = cpad(c); t.rows = (full ? Hp : rows) >> lvl; t.W = Wp >> lvl; t.esize = esize ? esize : act_esize;
        h->tensor_level.push_back(lvl); h->tensor_full.push_back(full ? 1 : 0);
        const bool pair = (flags == PTD_DN_3XTF32 || flags == PTD_DN_2XF16) && esize == 0;        // hi copy followed by the lo copy
        if (pair) t.lo_off = (t.floats() + 255) & ~(size_t)255;
        h->tensors.push_back(t);
        h->tensor_off.push_back(arena_bytes);
        arena_bytes += (((pair ? t.lo_off + t.floats() : t.floats())) * 4 + 1023) & ~(size_t)1023;
        return (int)h->tensors.size() - 1;
    };
    h->t_in16 = tnew(10, 0);
    const int encC[6] = {32, 43, 57, 76, 101, 101};
    int out1[6], mid[6], pooled[5];
    for (int l = 0; l < 6; ++l) {
        out1[l] = tnew(encC[l], l); mid[l] = tnew(encC[l], l);
        h->t_hidden[l][0] = tnew(encC[l], l); h->t_hidden[l][1] = tnew(encC[l], l);
        h->hidden_c[l] = encC[l];
        if (l < 5) pooled[l] = tnew(encC[l], l + 1);
    }
    const int decC[5] = {76, 57, 43, 32, 3};          // dec5 .. dec1 outputs, at levels 4 .. 0
    int dc1[5], dc2[5];
    for (int i = 0; i < 5; ++i) { dc1[i] = tnew(decC[i], 4 - i); dc2[i] = tnew(decC[i], 4 - i, i == 4 ? 4 : 0); }   // the denoised frame stays fp32
    h->t_final = dc2[4];
    if ((int)h->tensors.size() > DN_MAX_TENSORS) { ptd_set_error("ptd_dn_create: tensor table overflow"); return fail(PTD_ERR_STATE); }
    h->flags_off = arena_bytes;
    arena_bytes += h->tensors.size() * 2 * 32 + 1024;
    h->gflags_off = arena_bytes;
    arena_bytes += DN_MAX_RANKS * 32 + 1024;
    h->t_gather = (strip && h->repl_level >= 1 && h->repl_level <= 5) ? pooled[h->repl_level - 1] : -1;
    if (cudaMalloc((void**)&h->arena, arena_bytes) != cudaSuccess) { ptd_set_error("ptd_dn_create: cudaMalloc(%zu B activation arena) failed: %s", arena_bytes, cudaGetErrorString(cudaGetLastError())); return fail(PTD_ERR_CUDA); }
    h->arena_bytes = arena_bytes;
    cudaMemset(h->arena, 0, arena_bytes);
    for (size_t i = 0; i < h->tensors.size(); ++i) h->tensors[i].base = (float*)(h->arena + h->tensor_off[i]);

    // ---- wire the 28 convs (execution order of recurrent_autoencoder_model.py:129-140) ----
    // ids < 0 encode hidden states: -(10 + level) = "read parity", -(20 + level) = "write parity"
    auto HR = [](int l) { return -(10 + l); };
    auto HW = [](int l) { return -(20 + l); };
    auto resolve = [&](int id, int parity) -> int {
        if (id <= -20) return h->t_hidden[-id - 20][parity ^ 1];
        if (id <= -10) return h->t_hidden[-id - 10][parity];
        return id;
    };
    std::vector<DnLayerSpec> specs = dn_layer_specs();
    for (size_t li = 0; li < specs.size(); ++li) {
        DnLayer L;
        L.spec = specs[li];
        const DnLayerSpec& s = specs[li];
        const int lvl = s.level;
        L.H = ((strip && lvl >= h->repl_level) ? Hp : rows) >> lvl; L.W = Wp >> lvl;
        L.src1 = -1; L.pool = -1;
        if (s.kind == DN_L1) { L.src0 = lvl == 0 ? h->t_in16 : pooled[lvl - 1]; L.out = out1[lvl]; }
        else if (s.kind == DN_L2A) { L.src0 = out1[lvl]; L.src1 = HR(lvl); L.out = mid[lvl]; }
        else if (s.kind == DN_L2B) { L.src0 = mid[lvl]; L.out = HW(lvl); if (lvl < 5) L.pool = pooled[lvl]; }
        else if (s.kind == DN_DEC1) { const int i = 4 - lvl; L.src0 = i == 0 ? HW(5) : dc2[i - 1]; L.src1 = pooled[lvl]; L.out = dc1[i]; }   // dec5 reads the bottleneck output of THIS frame
        else { const int i = 4 - lvl; L.src0 = dc1[i]; L.out = dc2[i]; }
        const int c0p = cpad(s.cin0), c1p = s.cin1 ? cpad(s.cin1) : 0, coutp = cpad(s.cout);
        // ---- parameters: conv weight/bias + BN folded to scale/shift ----
        auto get = [&](const std::string& k, size_t want) -> const std::vector<float>* {
            auto it = sd.find(k);
            if (it == sd.end() || it->second.size() != want) { ptd_set_error("weight file: tensor '%s' missing or wrong size (want %zu)", k.c_str(), want); return nullptr; }
            return &it->second;
        };
        const int cin = s.cin0 + s.cin1, cinp = c0p + c1p;
        const std::vector<float>* w = get(s.conv_key + ".weight", (size_t)s.cout * cin * 9);
        const std::vector<float>* b = get(s.conv_key + ".bias", s.cout);
        const std::vector<float>* g = get(s.bn_key + ".weight", s.cout);
        const std::vector<float>* be = get(s.bn_key + ".bias", s.cout);
        const std::vector<float>* mu = get(s.bn_key + ".running_mean", s.cout);
        const std::vector<float>* var = get(s.bn_key + ".running_var", s.cout);
        if (!w || !b || !g || !be || !mu || !var) return fail(PTD_ERR_PARSE);
        std::vector<float> w9((size_t)9 * cinp * coutp, 0.f), scale(coutp, 0.f), shift(coutp, 0.f), bias(coutp, 0.f);
        auto cmap = [&](int c) { return c < s.cin0 ? c : c0p + (c - s.cin0); };      // position of real channel c in the padded concat
        for (int co = 0; co < s.cout; ++co)
            for (int c = 0; c < cin; ++c)
                for (int t = 0; t < 9; ++t)
                    w9[((size_t)t * cinp + cmap(c)) * coutp + co] = (*w)[((size_t)co * cin + c) * 9 + t];
        for (int co = 0; co < s.cout; ++co) {
            const float sc = (*g)[co] / sqrtf((*var)[co] + 1e-5f);                      // BatchNorm2d eval, eps 1e-5
            const float sh = (*be)[co] - (*mu)[co] * sc;
            scale[co] = sc; bias[co] = (*b)[co];
            shift[co] = s.lrelu_first ? sh : sh + sc * (*b)[co];
        }
        DALLOC(L.d_w9, w9.size()); DALLOC(L.d_scale, coutp); DALLOC(L.d_shift, coutp); DALLOC(L.d_bias, coutp);
        cudaMemcpy(L.d_w9, w9.data(), w9.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(L.d_scale, scale.data(), coutp * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(L.d_shift, shift.data(), coutp * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(L.d_bias, bias.data(), coutp * 4, cudaMemcpyHostToDevice);
        if (flags == PTD_DN_FP32_BATCH_STATS) {
            std::vector<float> gam(coutp, 0.f), bet(coutp, 0.f);
            for (int co = 0; co < s.cout; ++co) { gam[co] = (*g)[co]; bet[co] = (*be)[co]; }
            DALLOC(L.d_gamma, coutp); DALLOC(L.d_beta, coutp);
            cudaMemcpy(L.d_gamma, gam.data(), coutp * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(L.d_beta, bet.data(), coutp * 4, cudaMemcpyHostToDevice);
        }
        if (!dn_cuda_core_engine(flags)) {
            float* dd = nullptr;
            DALLOC(dd, 64);
            L.d_done = (uint32_t*)dd;
            for (int parity = 0; parity < 2; ++parity) {
                TcConvDesc d;
                d.src0 = h->tensors[resolve(L.src0, parity)];
                if (L.src1 != -1) d.src1 = h->tensors[resolve(L.src1, parity)];
                d.upsample = s.kind == DN_DEC1;
                d.out = h->tensors[resolve(L.out, parity)];
                if (L.pool != -1) d.pool_out = h->tensors[L.pool];
                d.lrelu_first = s.lrelu_first;
                d.scale = L.d_scale; d.shift = L.d_shift; d.bias = L.d_bias;
                d.round_out = li + 1 < specs.size();
                d.shared_wpack = parity ? L.tc[0].d_wpack : nullptr;
                // a strip-resident output computed from replicated (full-height) sources: decoder DN_REPL_LEVEL reads rows from row0 >> level on
                const int s0id = resolve(L.src0, parity);
                if (h->tensor_full[s0id] && !h->tensor_full[resolve(L.out, parity)]) d.src_yoff = row0 >> h->tensor_level[s0id];
                rc = tc_plan_create(d, w9, cinp, L.tc[parity], h->allocs);
                if (rc != PTD_OK) return fail(rc);
            }
        }
        h->layers.push_back(L);
        if (s.kind == DN_L2B && lvl < 5) h->pools.push_back({{h->t_hidden[lvl][0], h->t_hidden[lvl][1]}, pooled[lvl]});
    }
#undef DALLOC
    CUDA_TRY(cudaDeviceSynchronize());
    *out = h;
    return PTD_OK;
}

extern "C" ptd_status ptd_dn_create(const char* weights_path, int H, int W, int device, unsigned flags, ptd_dn** out) {
    return dn_create(weights_path, H, W, 0, 0, false, device, flags, out);
}
extern "C" ptd_status ptd_dn_create_strip(const char* weights_path, int H, int W, int row0, int rows, int device, unsigned flags, ptd_dn** out) {
    return dn_create(weights_path, H, W, row0, rows, true, device, flags, out);
}
// Rows of strip `index` of `nstrips`: the padded frame's 32-row groups split as evenly as possible (earlier strips get the extras).
extern "C" ptd_status ptd_dn_strip_partition(int H, int nstrips, int index, int* row0, int* rows) {
    const int groups = (H + 31) / 32;
    if (H <= 0 || nstrips < 1 || index < 0 || index >= nstrips || !row0 || !rows) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_strip_partition: bad argument");
    if (nstrips > groups) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_strip_partition: %d strips but the padded frame has only %d groups of 32 rows", nstrips, groups);
    const int base = groups / nstrips, extra = groups % nstrips;
    const int g0 = index * base + (index < extra ? index : extra), gn = base + (index < extra ? 1 : 0);
    *row0 = g0 * 32; *rows = gn * 32;
    return PTD_OK;
}
extern "C" int ptd_dn_strip_info_size(void) { return (int)sizeof(ptd_strip_info); }
extern "C" ptd_status ptd_dn_strip_export(ptd_dn* h, void* info_out, int capacity) {
    if (!h || !info_out || capacity < (int)sizeof(ptd_strip_info)) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_strip_export: need a buffer of %zu bytes", sizeof(ptd_strip_info));
    CUDA_TRY(cudaSetDevice(h->device));
    ptd_strip_info info;
    memset(&info, 0, sizeof info);
    cudaIpcMemHandle_t ipc;
    if (cudaIpcGetMemHandle(&ipc, h->arena) == cudaSuccess) memcpy(info.ipc, &ipc, sizeof ipc);
    else cudaGetLastError();                                  // same-process connections do not need it
    info.arena = (unsigned long long)(uintptr_t)h->arena;
    info.pid_tag = (int)getpid(); info.device = h->device; info.row0 = h->row0; info.rows = h->rows; info.Hp = h->Hp; info.Wp = h->Wp;
    info.ntensors = (int)h->tensors.size();
    for (size_t i = 0; i < h->tensors.size(); ++i) { info.tensor_off[i] = h->tensor_off[i]; info.tensor_rows[i] = h->tensors[i].rows; }
    info.flags_off = h->flags_off;
    info.gflags_off = h->gflags_off;
    info.reserved = h->repl_level;
    memcpy(info_out, &info, sizeof info);
    return PTD_OK;
}
// infos: the exported blobs of ALL strips of the frame in strip order (top strip first), back to back; my_rank: this handle's
// position.  Same process: the pointers are used directly (peer access is enabled when the devices differ); other process: the
// arenas are opened through CUDA IPC.  The strips above / below receive halo rows, every strip receives the gathered level.
extern "C" ptd_status ptd_dn_strip_connect(ptd_dn* h, const void* infos, int nranks, int my_rank) {
    if (!h || !infos || nranks < 1 || nranks > DN_MAX_RANKS || my_rank < 0 || my_rank >= nranks) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_strip_connect: bad argument (at most %d strips)", DN_MAX_RANKS);
    CUDA_TRY(cudaSetDevice(h->device));
    const ptd_strip_info* in = (const ptd_strip_info*)infos;
    int row = 0;
    for (int r = 0; r < nranks; ++r) {
        ptd_strip_info info;
        memcpy(&info, &in[r], sizeof info);
        if (info.ntensors != (int)h->tensors.size() || info.Hp != h->Hp || info.Wp != h->Wp) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_strip_connect: strip %d was built for another frame size", r);
        if (info.reserved != h->repl_level) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_strip_connect: strip %d replicates from level %d, this one from %d (PTD_DN_REPL_LEVEL must agree on all ranks)", r, info.reserved, h->repl_level);
        if (info.row0 != row) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_strip_connect: strip %d starts at row %d, expected %d", r, info.row0, row);
        row += info.rows;
        if (r == my_rank && (info.row0 != h->row0 || info.rows != h->rows)) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_strip_connect: blob %d is not this handle's", r);
    }
    if (row != h->Hp) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_strip_connect: the strips cover %d of %d padded rows", row, h->Hp);
    for (int r = 0; r < DN_MAX_RANKS; ++r) {
        if (h->peer_arena[r] && h->peer_ipc[r]) cudaIpcCloseMemHandle(h->peer_arena[r]);
        h->peer_arena[r] = nullptr; h->peer_ipc[r] = false;
    }
    for (int r = 0; r < nranks; ++r) {
        if (r == my_rank) continue;
        ptd_strip_info info;
        memcpy(&info, &in[r], sizeof info);
        if (info.pid_tag == (int)getpid()) {
            if (info.device != h->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(info.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ptd_set_error("ptd_dn_strip_connect: no peer access %d -> %d: %s", h->device, info.device, cudaGetErrorString(e)); return PTD_ERR_CUDA; }
                cudaGetLastError();
            }
            h->peer_arena[r] = (unsigned char*)(uintptr_t)info.arena;
        } else {
            cudaIpcMemHandle_t ipc;
            memcpy(&ipc, info.ipc, sizeof ipc);
            void* q = nullptr;
            CUDA_TRY(cudaIpcOpenMemHandle(&q, ipc, cudaIpcMemLazyEnablePeerAccess));
            h->peer_arena[r] = (unsigned char*)q; h->peer_ipc[r] = true;
        }
        h->peer_info[r] = info;
    }
    h->rank = my_rank; h->nranks = nranks;
    h->has_peer[0] = my_rank > 0; h->has_peer[1] = my_rank + 1 < nranks;
    return PTD_OK;
}

// dir: 0 = the strip above (rank - 1), 1 = the strip below (rank + 1)
static DnTensor peer_tensor(const ptd_dn* h, int dir, int id) {
    const int r = h->rank + (dir == 0 ? -1 : 1);
    DnTensor t = h->tensors[id];
    t.base = (float*)(h->peer_arena[r] + h->peer_info[r].tensor_off[id]);
    t.rows = h->peer_info[r].tensor_rows[id];
    if (t.lo_off) t.lo_off = (t.floats() + 255) & ~(size_t)255;       // the neighbour's hi -> lo distance follows ITS row count
    return t;
}
static uint32_t* peer_flag(const ptd_dn* h, int dir, int id, int from) {
    const int r = h->rank + (dir == 0 ? -1 : 1);
    return (uint32_t*)(h->peer_arena[r] + h->peer_info[r].flags_off + ((size_t)id * 2 + from) * 32);
}

// Launches layers [first, last) of forward(x, j) on `st`.  ptd_dn_forward runs them all; a same-GPU strip group (tests) interleaves.
static ptd_status dn_run(ptd_dn* h, const float* gbuf, float* rgb, int reset_hidden, cudaStream_t st, int first, int last) {
    CUDA_TRY(cudaSetDevice(h->device));
    const int nl = (int)h->layers.size();
    auto mark = [&](const char* name) {                                 // event after the launch just issued (and one before the first)
        if (!h->profiling) return;
        const int nmark = (int)h->launch_names.size() + (name ? 1 : 0);
        if ((int)h->events.size() <= nmark) { cudaEvent_t e; cudaEventCreate(&e); h->events.push_back(e); }
        cudaEventRecord(h->events[name ? nmark : 0], st);
        if (name) h->launch_names.push_back(name);
    };
    const bool tf32 = h->flags == PTD_DN_TF32;               // fp32 storage, operands rounded to tf32 at the producer
    const bool tensor = !dn_cuda_core_engine(h->flags);
    const bool batch_stats = h->flags == PTD_DN_FP32_BATCH_STATS;
    if (first <= -1) {                                                  // stage -1: start of the frame
        h->epoch += 1;
        h->launches = 0;
        if (h->profiling) h->launch_names.clear();
        if (reset_hidden)                                               // forward(x, j == 0): model.py:121-128
            for (int l = 0; l < 6; ++l) {
                const DnTensor& t = h->tensors[h->t_hidden[l][h->parity]];
                CUDA_TRY(cudaMemsetAsync(t.base, 0, (t.lo_off ? t.lo_off + t.floats() : t.floats()) * 4, st));
            }
        const DnTensor& in = h->tensors[h->t_in16];
        const size_t n = (size_t)(in.rows + 2) * in.W;
        mark(nullptr);
        PackLink link;
        memset(&link, 0, sizeof link);
        if (h->has_peer[0] || h->has_peer[1]) {
            link.done = h->d_pack_done; link.epoch = h->epoch;
            if (h->has_peer[0]) { link.up = peer_tensor(h, 0, h->t_in16); link.sig_up = peer_flag(h, 0, h->t_in16, 1); }
            if (h->has_peer[1]) { link.down = peer_tensor(h, 1, h->t_in16); link.sig_down = peer_flag(h, 1, h->t_in16, 0); }
        }
        pack_gbuffer<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gbuf, h->H, h->W, h->row0, in, tf32, link);
        ++h->launches;
        mark("pack_gbuffer");
    }
    for (int li = first < 0 ? 0 : first; li < last && li < nl; ++li) {
        DnLayer& L = h->layers[li];
        auto res = [&](int id) -> int {
            if (id <= -20) return h->t_hidden[-id - 20][h->parity ^ 1];
            if (id <= -10) return h->t_hidden[-id - 10][h->parity];
            return id;
        };
        const int s0 = res(L.src0), s1 = L.src1 == -1 ? -1 : res(L.src1), o = res(L.out);
        if (tensor) {
            TcConvPlan& plan = L.tc[h->parity];
            TcStripLink& k = plan.p.link;
            memset(&k, 0, sizeof k);
            if (h->nranks > 1) {
                k.epoch = h->epoch;
                const int srcs[2] = {s0, s1};
                for (int i = 0; i < 2; ++i) {
                    if (srcs[i] < 0) continue;
                    if (h->tensor_full[srcs[i]]) {
                        // replicated level: complete on this GPU - except the gathered tensor, whose rows arrive from every other strip
                        if (srcs[i] == h->t_gather && L.spec.kind == DN_L1)
                            for (int r = 0; r < h->nranks; ++r)
                                if (r != h->rank) k.gather_wait[r] = (const uint32_t*)(h->arena + h->gflags_off + (size_t)r * 32);
                        continue;
                    }
                    // a hidden state read by layer2's first conv was produced by the PREVIOUS frame (or just zeroed)
                    const bool prev = L.spec.kind == DN_L2A && i == 1;
                    if (prev && reset_hidden) continue;
                    for (int d = 0; d < 2; ++d)
                        if (h->has_peer[d]) { k.wait[2 * i + d] = h->flag(srcs[i], d); k.wait_epoch[2 * i + d] = prev ? h->epoch - 1 : h->epoch; }
                }
                if (!h->tensor_full[o]) {
                    k.done = L.d_done;
                    for (int d = 0; d < 2; ++d) {
                        if (!h->has_peer[d]) continue;
                        // our first row -> the UP neighbour's bottom apron, flagged there as "from down" (1); last row -> DOWN neighbour, "from up" (0)
                        (d == 0 ? k.out_up : k.out_down) = peer_tensor(h, d, o);
                        k.sig[d] = peer_flag(h, d, o, d == 0 ? 1 : 0);
                        if (L.pool != -1 && !h->tensor_full[L.pool]) {
                            (d == 0 ? k.pool_up : k.pool_down) = peer_tensor(h, d, L.pool);
                            k.sig[2 + d] = peer_flag(h, d, L.pool, d == 0 ? 1 : 0);
                        }
                    }
                    if (L.pool != -1 && h->tensor_full[L.pool]) {           // the gather: our pooled rows go into every strip's full-height tensor
                        k.pool_yoff = h->row0 >> h->tensor_level[L.pool];
                        for (int r = 0; r < h->nranks; ++r) {
                            if (r == h->rank) continue;
                            k.gather_base[r] = (float*)(h->peer_arena[r] + h->peer_info[r].tensor_off[L.pool]);
                            k.gather_sig[r] = (uint32_t*)(h->peer_arena[r] + h->peer_info[r].gflags_off + (size_t)h->rank * 32);
                        }
                    }
                }
            } else if (L.pool != -1 && h->tensor_full[L.pool] && !h->tensor_full[o]) {
                k.pool_yoff = h->row0 >> h->tensor_level[L.pool];           // a single strip that is not the whole frame cannot exist; kept for symmetry
            }
            plan.p.linked = h->nranks > 1 ? 1 : 0;
            plan.p.pdl = (h->pdl && li > 0 && !h->profiling) ? 1 : 0;       // the first conv follows pack_gbuffer, which does not trigger early
            ptd_status rc = tc_conv_launch(plan, st, &h->launches, nullptr, h->sm_limit);
            if (rc != PTD_OK) return rc;
            mark(L.spec.name.c_str());
        } else {
            FpConvArgs a;
            a.src0 = h->tensors[s0]; a.src1 = s1 >= 0 ? h->tensors[s1] : DnTensor(); a.upsample = L.spec.kind == DN_DEC1;
            a.H = L.H; a.W = L.W; a.w = L.d_w9; a.scale = L.d_scale; a.shift = L.d_shift; a.bias = L.d_bias;
            a.order = L.spec.lrelu_first ? 1 : 0; a.out = h->tensors[o];
            dim3 grid((L.W + FC_TW - 1) / FC_TW, (L.H + FC_TH - 1) / FC_TH, (a.out.cp + FC_BN - 1) / FC_BN);
            if (batch_stats) {
                a.order = L.spec.lrelu_first ? 3 : 2;
                conv3x3_fp32<1><<<grid, 128, 0, st>>>(a);
                CUDA_TRY(cudaMemsetAsync(h->d_bn_stats, 0, sizeof(double) * 2 * 128, st));
                const size_t npx = (size_t)a.out.rows * a.out.W;
                dim3 sgrid((unsigned)std::min<size_t>((npx + 255) / 256, 592), a.out.cp / 4);
                bn_batch_stats<<<sgrid, 256, 0, st>>>(a.out, h->d_bn_stats);
                bn_batch_apply<<<sgrid, 256, 0, st>>>(a.out, h->d_bn_stats, L.d_gamma, L.d_beta, L.spec.lrelu_first ? 0 : 1);
                h->launches += 3;
            } else {
                conv3x3_fp32<0><<<grid, 128, 0, st>>>(a);
                ++h->launches;
            }
            mark(L.spec.name.c_str());
            if (L.pool != -1) {
                const DnTensor& po = h->tensors[L.pool];
                const size_t n = (size_t)(po.cp / 4) * po.rows * po.W;
                maxpool2_chw4<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->tensors[o], po);
                ++h->launches;
                mark("maxpool2");
            }
        }
    }
    if (last > nl) {                                                    // stage nl: end of the frame
        const int r0 = h->row0, r1 = std::min(h->row0 + h->rows, h->H);
        if (r1 > r0) {
            const size_t n = (size_t)(r1 - r0) * h->W;
            unpack_rgb<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->tensors[h->t_final], h->H, h->W, r0, r1 - r0, rgb);
            ++h->launches;
            mark("unpack_rgb");
        }
        h->parity ^= 1;
        if (h->profiling) h->timed_launches = (int)h->launch_names.size();
    }
    CUDA_TRY(cudaGetLastError());
    return PTD_OK;
}

// While frames submitted with ptd_frame_submit are in flight they own the recurrent state, the parity and the arena on streams of their
// own; any other entry point would race with them, so it reports PTD_ERR_STATE instead.
#define DN_NOT_INFLIGHT(h, what) do { if ((h)->inflight > 0) PTD_FAIL(PTD_ERR_STATE, what ": %d frame(s) submitted with ptd_frame_submit are still in flight - call ptd_frame_wait first", (h)->inflight); } while (0)
ptd_status ptd_dn_forward_frame(ptd_dn* h, const float* gbuf, float* rgb, int reset_hidden, void* stream) {       // ptd_frame_* only (ptd_pt.cu)
    return dn_run(h, gbuf, rgb, reset_hidden, (cudaStream_t)stream, -1, (int)h->layers.size() + 1);
}
void ptd_dn_set_sm_limit(ptd_dn* h, int sms) { h->sm_limit = sms; }
void ptd_dn_mark_inflight(ptd_dn* h, int delta) { h->inflight += delta; if (h->inflight < 0) h->inflight = 0; }
extern "C" ptd_status ptd_dn_forward(ptd_dn* h, const float* gbuf, float* rgb, int reset_hidden, void* stream_) {
    if (!h || !gbuf || !rgb) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_forward: null argument");
    DN_NOT_INFLIGHT(h, "ptd_dn_forward");
    return dn_run(h, gbuf, rgb, reset_hidden, (cudaStream_t)stream_, -1, (int)h->layers.size() + 1);
}

// Same-process strip group (one or several devices): the layers of all strips are issued layer by layer, so that on a
// single GPU - where the strips' kernels cannot run concurrently - every flag a kernel waits for has already been raised.
extern "C" ptd_status ptd_dn_forward_group(ptd_dn** hs, int n, const float* const* gbufs, float* const* rgbs, int reset_hidden, void* const* streams) {
    if (!hs || n < 1 || !gbufs || !rgbs) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_forward_group: bad argument");
    const int nl = (int)hs[0]->layers.size();
    for (int stage = -1; stage <= nl; ++stage)
        for (int i = 0; i < n; ++i) {
            ptd_status rc = dn_run(hs[i], gbufs[i], rgbs[i], reset_hidden, streams ? (cudaStream_t)streams[i] : nullptr, stage, stage == -1 ? 0 : stage + 1);
            if (rc != PTD_OK) return rc;
        }
    return PTD_OK;
}

extern "C" ptd_status ptd_dn_forward_host(ptd_dn* h, const float* gbuf_host, float* rgb_host, int reset_hidden) {
    if (!h || !gbuf_host || !rgb_host) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_forward_host: null argument");
    if (h->strip) PTD_FAIL(PTD_ERR_UNSUPPORTED, "ptd_dn_forward_host: a row-strip handle takes device pointers (ptd_dn_forward)");
    DN_NOT_INFLIGHT(h, "ptd_dn_forward_host");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t P = (size_t)h->H * h->W;
    CUDA_TRY(cudaMemcpy(h->d_gbuf, gbuf_host, P * 40, cudaMemcpyHostToDevice));          // main.cpp:104-105
    ptd_status rc = ptd_dn_forward(h, h->d_gbuf, h->d_rgb, reset_hidden, nullptr);
    if (rc != PTD_OK) return rc;
    CUDA_TRY(cudaMemcpy(rgb_host, h->d_rgb, P * 12, cudaMemcpyDeviceToHost));            // main.cpp:91 (.to(kCPU))
    return PTD_OK;
}

void ptd_dn_describe(const ptd_dn* h, int* device, int* H, int* W, int* strip) {
    *device = h->device; *H = h->H; *W = h->W; *strip = h->strip ? 1 : 0;
}

extern "C" ptd_status ptd_dn_padded_size(const ptd_dn* h, int* Hp, int* Wp) {
    if (!h) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_padded_size: null handle");
    if (Hp) *Hp = h->Hp;
    if (Wp) *Wp = h->Wp;
    return PTD_OK;
}

extern "C" ptd_status ptd_dn_dump_hidden(ptd_dn* h, int level, float* host, size_t cap, int* C, int* H, int* W) {
    if (!h || !host || level < 0 || level > 5) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_dump_hidden: bad argument");
    DN_NOT_INFLIGHT(h, "ptd_dn_dump_hidden");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());                                  // the last forward may have run on a non-blocking stream
    const DnTensor& t = h->tensors[h->t_hidden[level][h->parity]];       // the state the NEXT forward will read
    const int c = h->hidden_c[level], hh = t.rows, ww = t.W;
    const size_t n = (size_t)c * hh * ww;
    if (cap < n) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_dump_hidden: capacity %zu < %zu", cap, n);
    float* tmp = nullptr;
    CUDA_TRY(cudaMalloc((void**)&tmp, n * 4));
    chw4_to_nchw<<<(unsigned)((n + 255) / 256), 256>>>(t, c, tmp);
    cudaError_t e = cudaMemcpy(host, tmp, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    CUDA_TRY(e);
    if (C) *C = c;
    if (H) *H = hh;
    if (W) *W = ww;
    return PTD_OK;
}
extern "C" int ptd_dn_launches_per_forward(const ptd_dn* h) { return h ? h->launches : 0; }
extern "C" ptd_status ptd_dn_profile(ptd_dn* h, int enable) {
    if (!h) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_profile: null handle");
    h->profiling = enable != 0;
    h->timed_launches = 0;
    return PTD_OK;
}
extern "C" ptd_status ptd_dn_launch_times(ptd_dn* h, float* ms, int capacity, int* n) {
    if (!h || !ms || !n) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_launch_times: null argument");
    if (h->timed_launches <= 0) PTD_FAIL(PTD_ERR_STATE, "ptd_dn_launch_times: no profiled forward has run");
    if (capacity < h->timed_launches) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_launch_times: capacity %d < %d", capacity, h->timed_launches);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventSynchronize(h->events[h->timed_launches]));
    for (int i = 0; i < h->timed_launches; ++i) CUDA_TRY(cudaEventElapsedTime(&ms[i], h->events[i], h->events[i + 1]));
    *n = h->timed_launches;
    return PTD_OK;
}
extern "C" const char* ptd_dn_launch_name(const ptd_dn* h, int i) {
    return (h && i >= 0 && i < (int)h->launch_names.size()) ? h->launch_names[i].c_str() : "";
}
