// HP-2: forward pass of the recurrent denoising autoencoder, B200-native.
//
// Replaces network_prediction_faster_version() / torch::jit Module.forward (Inference/src/main.cpp:101-118) for the
// network defined in training/recurrent_autoencoder_model.py:8-142 - 28 x (conv3x3 pad 1 + BatchNorm(eval) + LeakyReLU 0.1),
// 5 x MaxPool2, 5 x nearest Upsample x2, 11 x channel concat, 6 recurrent hidden states - with eval-mode BN and carried
// hidden state (SURVEY.md decisions D1, D3).
//
// Data layout in HBM: every activation is NHWC fp32 with the channel count padded to a multiple of 16 (10->16, 32, 43->48,
// 57->64, 76->80, 101->112, 3->16; pad channels are kept at exactly 0), so one pixel's channels are one contiguous,
// 64-byte-aligned run: coalesced for the CUDA-core path and a legal TMA box for the tensor-core path.  Concats and the
// decoder's upsample are never materialised (the convs read two sources / index at half resolution); conv bias, BN and
// LeakyReLU are folded into the conv epilogue.  Two conv engines share the buffers and the layer graph:
//   PTD_DN_FP32  conv3x3_fp32   - implicit-GEMM on CUDA cores (FFMA), strict-parity path;
//   PTD_DN_TF32  dn_conv_tc.cuh - TMA-staged tiles + tcgen05.mma kind::tf32 with TMEM accumulators.
#include <cuda_runtime.h>
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
    const float* src0; const float* src1;   // NHWC, padded channel counts c0p / c1p (src1 may be null)
    int c0p, c1p;
    int upsample;                           // sources live at (H/2, W/2): nearest x2 (model.py:40)
    int H, W;                               // output (= conv input) resolution
    const float* w;                         // [9][c0p + c1p][coutp]
    const float* scale; const float* shift; const float* bias;   // [coutp]
    int coutp; int order;                   // 0: lrelu(scale*acc + shift)   1: scale*lrelu(acc + bias) + shift
    float* out;                             // NHWC [H][W][coutp]
};

__global__ void __launch_bounds__(128) conv3x3_fp32(const FpConvArgs a) {
    __shared__ float s_in[FC_CK * FC_PS];
    __shared__ __align__(16) float s_w[9 * FC_CK * FC_BN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * FC_TW, y0 = blockIdx.y * FC_TH, co0 = blockIdx.z * FC_BN;
    const int col = lane & 15, rbase = (lane >> 4) * 4;     // thread's pixels: rows rbase..rbase+3 of the tile, column col
    const int cg = warp * 8;                                  // thread's 8 output channels inside the 32-wide tile
    const int cin = a.c0p + a.c1p;
    const int sH = a.upsample ? a.H >> 1 : a.H, sW = a.upsample ? a.W >> 1 : a.W;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < cin; c0 += FC_CK) {
        const float* src = c0 < a.c0p ? a.src0 : a.src1;
        const int cs = c0 < a.c0p ? c0 : c0 - a.c0p, cp = c0 < a.c0p ? a.c0p : a.c1p;
        __syncthreads();
        // halo tile: 10 x 18 pixels x 8 channels (two float4 per pixel), zero outside the image (padding = 1)
        for (int i = tid; i < 10 * 18 * 2; i += 128) {
            const int half = i & 1, pix = i >> 1, r = pix / 18, c = pix % 18;
            const int gy = y0 + r - 1, gx = x0 + c - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
                const int sy = a.upsample ? gy >> 1 : gy, sx = a.upsample ? gx >> 1 : gx;
                v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)sy * sW + sx) * cp + cs + half * 4));
            }
            float* d = s_in + (half * 4) * FC_PS + r * FC_RS + c;
            d[0] = v.x; d[FC_PS] = v.y; d[2 * FC_PS] = v.z; d[3 * FC_PS] = v.w;
        }
        // weights of this chunk: [9][8][32]
        for (int i = tid; i < 9 * FC_CK * FC_BN / 4; i += 128) {
            const int j4 = i % (FC_BN / 4), c = (i / (FC_BN / 4)) % FC_CK, tap = i / (FC_BN / 4 * FC_CK);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (co0 + j4 * 4 < a.coutp) v = __ldg(reinterpret_cast<const float4*>(a.w + ((size_t)tap * cin + c0 + c) * a.coutp + co0 + j4 * 4));
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
    if (co >= a.coutp) return;
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
            if (a.order == 0) { float v = fmaf(acc[i][j], sc[j], sh[j]); o[j] = v > 0.f ? v : 0.1f * v; }
            else { float v = acc[i][j] + bi[j]; v = v > 0.f ? v : 0.1f * v; o[j] = fmaf(v, sc[j], sh[j]); }
        }
        float4* dst = reinterpret_cast<float4*>(a.out + ((size_t)gy * a.W + gx) * a.coutp + co);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// MaxPool2d(2) on NHWC (model.py:19): one thread = one output pixel x 4 channels
__global__ void maxpool2_nhwc(const float* __restrict__ in, float* __restrict__ out, int Ho, int Wo, int cp) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = cp / 4;
    if (i >= (size_t)Ho * Wo * c4) return;
    const int c = (int)(i % c4), x = (int)((i / c4) % Wo), y = (int)(i / ((size_t)c4 * Wo));
    const float4* p = reinterpret_cast<const float4*>(in) + ((size_t)(2 * y) * (2 * Wo) + 2 * x) * c4 + c;
    const float4 a = p[0], b = p[c4], d = p[(size_t)2 * Wo * c4], e = p[(size_t)2 * Wo * c4 + c4];
    reinterpret_cast<float4*>(out)[i] = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                                                    fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
}
// planar G-buffer [10][H][W] rows [row0, row0+rows) -> NHWC16 [rows_p][Wp][16], zero padded (decision D3)
__global__ void pack_gbuffer(const float* __restrict__ g, int H, int W, int Hp, int Wp, float* __restrict__ out, int round_tf32) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)Hp * Wp) return;
    const int x = (int)(i % Wp), y = (int)(i / Wp);
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.f;
    if (x < W && y < H) {
#pragma unroll
        for (int c = 0; c < 10; ++c) v[c] = g[(size_t)c * H * W + (size_t)y * W + x];
        if (round_tf32) {
#pragma unroll
            for (int c = 0; c < 10; ++c) v[c] = tc::round_tf32(v[c]);
        }
    }
    float4* d = reinterpret_cast<float4*>(out + i * 16);
    d[0] = make_float4(v[0], v[1], v[2], v[3]); d[1] = make_float4(v[4], v[5], v[6], v[7]);
    d[2] = make_float4(v[8], v[9], v[10], v[11]); d[3] = make_float4(v[12], v[13], v[14], v[15]);
}
// NHWC16 -> planar [3][H][W], cropped
__global__ void unpack_rgb(const float* __restrict__ in, int H, int W, int Wp, float* __restrict__ rgb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)H * W) return;
    const int x = (int)(i % W), y = (int)(i / W);
    const float4 v = *reinterpret_cast<const float4*>(in + ((size_t)y * Wp + x) * 16);
    rgb[i] = v.x; rgb[(size_t)H * W + i] = v.y; rgb[(size_t)2 * H * W + i] = v.z;
}
// NHWC padded -> NCHW (hidden-state parity tap)
__global__ void nhwc_to_nchw(const float* __restrict__ in, int C, int cp, int H, int W, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)C * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)(i / ((size_t)W * H));
    out[i] = in[((size_t)y * W + x) * cp + c];
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
struct DnLayer {
    DnLayerSpec spec;
    int H, W;                       // output resolution
    const float* src0; const float* src1; int c0p, c1p;
    float* out; int coutp;
    float* d_w9 = nullptr;          // [9][cin_p][coutp]  (CUDA-core engine)
    float* d_scale = nullptr; float* d_shift = nullptr; float* d_bias = nullptr;
    TcConvPlan tc;                  // tensor-core engine plan (tensor maps, packed weights)
};

struct ptd_dn {
    int device = 0; unsigned flags = 0;
    int H = 0, W = 0, Hp = 0, Wp = 0;
    std::vector<DnLayer> layers;
    std::vector<void*> allocs;
    float* d_in16 = nullptr;        // packed input
    float* d_gbuf = nullptr; float* d_rgb = nullptr;   // staging for the host-pointer entry point
    float* hidden[6] = {nullptr}; int hidden_c[6] = {0}; size_t hidden_bytes[6] = {0};
    struct Pool { const float* in; float* out; int Ho, Wo, cp; };
    std::vector<Pool> pools;        // pools[k] follows layer 3k+2
    float* d_final = nullptr;
    int launches = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> events;          // events[i], events[i+1] bracket launch i
    std::vector<std::string> launch_names;
    int timed_launches = 0;
};

extern "C" void ptd_dn_destroy(ptd_dn* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto& L : h->layers) tc_plan_destroy(L.tc);
    for (void* p : h->allocs) cudaFree(p);
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    delete h;
}

static inline int cpad(int c) { return (c + 15) & ~15; }

extern "C" ptd_status ptd_dn_create(const char* weights_path, int H, int W, int device, unsigned flags, ptd_dn** out) {
    if (!weights_path || !out || H <= 0 || W <= 0) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_create: bad argument");
    *out = nullptr;
    if (ptd_device_count() <= device || device < 0) PTD_FAIL(PTD_ERR_CUDA, "ptd_dn_create: CUDA device %d not available (no CPU fallback exists)", device);
    if (flags != PTD_DN_FP32 && flags != PTD_DN_TF32) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_create: unknown flags %u", flags);
    std::map<std::string, std::vector<float>> sd;
    ptd_status rc = read_ptdw(weights_path, sd);
    if (rc != PTD_OK) return rc;
    CUDA_TRY(cudaSetDevice(device));
    ptd_dn* h = new ptd_dn();
    h->device = device; h->flags = flags; h->H = H; h->W = W;
    h->Hp = (H + 31) / 32 * 32; h->Wp = (W + 31) / 32 * 32;
    auto fail = [&](ptd_status code) { ptd_dn_destroy(h); return code; };
    auto dalloc = [&](size_t floats) -> float* {
        void* p = nullptr;
        if (cudaMalloc(&p, floats * 4 + 256) != cudaSuccess) { ptd_set_error("ptd_dn_create: cudaMalloc(%zu B) failed: %s", floats * 4, cudaGetErrorString(cudaGetLastError())); return nullptr; }
        cudaMemset(p, 0, floats * 4 + 256);
        h->allocs.push_back(p);
        return (float*)p;
    };
#define DALLOC(var, floats) do { (var) = dalloc(floats); if (!(var)) return fail(PTD_ERR_CUDA); } while (0)
    const size_t px0 = (size_t)h->Hp * h->Wp;
    DALLOC(h->d_in16, px0 * 16);
    DALLOC(h->d_gbuf, (size_t)10 * H * W);
    DALLOC(h->d_rgb, (size_t)3 * H * W);

    std::vector<DnLayerSpec> specs = dn_layer_specs();
    // activation buffers
    const int encC[6] = {32, 43, 57, 76, 101, 101};
    float* out1[6]; float* mid[6]; float* pooled[5];
    for (int l = 0; l < 6; ++l) {
        const size_t px = px0 >> (2 * l);
        const int cp = cpad(encC[l]);
        DALLOC(out1[l], px * cp); DALLOC(mid[l], px * cp); DALLOC(h->hidden[l], px * cp);
        h->hidden_c[l] = encC[l]; h->hidden_bytes[l] = px * cp * 4;
        if (l < 5) DALLOC(pooled[l], (px >> 2) * cp);
    }
    const int decC[5] = {76, 57, 43, 32, 3};          // dec5 .. dec1 outputs, at levels 4 .. 0
    float* dc1[5]; float* dc2[5];
    for (int i = 0; i < 5; ++i) {
        const int lvl = 4 - i;
        const size_t px = px0 >> (2 * lvl);
        DALLOC(dc1[i], px * cpad(decC[i])); DALLOC(dc2[i], px * cpad(decC[i]));
    }
    h->d_final = dc2[4];

    // wire the 28 convs (execution order of recurrent_autoencoder_model.py:129-140)
    for (size_t li = 0; li < specs.size(); ++li) {
        DnLayer L;
        L.spec = specs[li];
        const DnLayerSpec& s = specs[li];
        const int lvl = s.level;
        L.H = h->Hp >> lvl; L.W = h->Wp >> lvl;
        L.src1 = nullptr; L.c1p = 0;
        if (s.kind == DN_L1) {
            L.src0 = lvl == 0 ? h->d_in16 : pooled[lvl - 1]; L.c0p = cpad(s.cin0);
            L.out = out1[lvl];
        } else if (s.kind == DN_L2A) {
            L.src0 = out1[lvl]; L.c0p = cpad(s.cin0); L.src1 = h->hidden[lvl]; L.c1p = cpad(s.cin1);
            L.out = mid[lvl];
        } else if (s.kind == DN_L2B) {
            L.src0 = mid[lvl]; L.c0p = cpad(s.cin0);
            L.out = h->hidden[lvl];
        } else if (s.kind == DN_DEC1) {
            const int i = 4 - lvl;                     // dec index: level 4 -> dec5 (i = 0)
            L.src0 = i == 0 ? h->hidden[5] : dc2[i - 1]; L.c0p = cpad(s.cin0);
            L.src1 = pooled[lvl]; L.c1p = cpad(s.cin1);
            L.out = dc1[i];
        } else {
            const int i = 4 - lvl;
            L.src0 = dc1[i]; L.c0p = cpad(s.cin0);
            L.out = dc2[i];
        }
        L.coutp = cpad(s.cout);
        // ---- parameters: conv weight/bias + BN folded to scale/shift ----
        auto get = [&](const std::string& k, size_t want) -> const std::vector<float>* {
            auto it = sd.find(k);
            if (it == sd.end() || it->second.size() != want) { ptd_set_error("weight file: tensor '%s' missing or wrong size (want %zu)", k.c_str(), want); return nullptr; }
            return &it->second;
        };
        const int cin = s.cin0 + s.cin1, cinp = L.c0p + L.c1p;
        const std::vector<float>* w = get(s.conv_key + ".weight", (size_t)s.cout * cin * 9);
        const std::vector<float>* b = get(s.conv_key + ".bias", s.cout);
        const std::vector<float>* g = get(s.bn_key + ".weight", s.cout);
        const std::vector<float>* be = get(s.bn_key + ".bias", s.cout);
        const std::vector<float>* mu = get(s.bn_key + ".running_mean", s.cout);
        const std::vector<float>* var = get(s.bn_key + ".running_var", s.cout);
        if (!w || !b || !g || !be || !mu || !var) return fail(PTD_ERR_PARSE);
        std::vector<float> w9((size_t)9 * cinp * L.coutp, 0.f), scale(L.coutp, 0.f), shift(L.coutp, 0.f), bias(L.coutp, 0.f);
        auto cmap = [&](int c) { return c < s.cin0 ? c : L.c0p + (c - s.cin0); };      // position of real channel c in the padded concat
        for (int co = 0; co < s.cout; ++co)
            for (int c = 0; c < cin; ++c)
                for (int t = 0; t < 9; ++t)
                    w9[((size_t)t * cinp + cmap(c)) * L.coutp + co] = (*w)[((size_t)co * cin + c) * 9 + t];
        for (int co = 0; co < s.cout; ++co) {
            const float sc = (*g)[co] / sqrtf((*var)[co] + 1e-5f);                      // BatchNorm2d eval, eps 1e-5
            const float sh = (*be)[co] - (*mu)[co] * sc;
            scale[co] = sc; bias[co] = (*b)[co];
            shift[co] = s.lrelu_first ? sh : sh + sc * (*b)[co];
        }
        DALLOC(L.d_w9, w9.size()); DALLOC(L.d_scale, L.coutp); DALLOC(L.d_shift, L.coutp); DALLOC(L.d_bias, L.coutp);
        cudaMemcpy(L.d_w9, w9.data(), w9.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(L.d_scale, scale.data(), L.coutp * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(L.d_shift, shift.data(), L.coutp * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(L.d_bias, bias.data(), L.coutp * 4, cudaMemcpyHostToDevice);
        if (flags == PTD_DN_TF32) {
            TcConvDesc d;
            d.src0 = L.src0; d.src1 = L.src1; d.c0p = L.c0p; d.c1p = L.c1p; d.upsample = s.kind == DN_DEC1;
            d.H = L.H; d.W = L.W; d.coutp = L.coutp; d.out = L.out; d.lrelu_first = s.lrelu_first;
            d.scale = L.d_scale; d.shift = L.d_shift; d.bias = L.d_bias;
            d.pool_out = (s.kind == DN_L2B && lvl < 5) ? pooled[lvl] : nullptr;
            d.round_out = li + 1 < specs.size();
            rc = tc_plan_create(d, w9, cinp, L.tc, h->allocs);
            if (rc != PTD_OK) return fail(rc);
        }
        h->layers.push_back(L);
        if (s.kind == DN_L2B && lvl < 5) h->pools.push_back({h->hidden[lvl], pooled[lvl], L.H / 2, L.W / 2, L.coutp});
    }
#undef DALLOC
    CUDA_TRY(cudaDeviceSynchronize());
    *out = h;
    return PTD_OK;
}

extern "C" ptd_status ptd_dn_create_strip(const char*, int, int, int, int, int, unsigned, ptd_halo_fn, void*, ptd_dn** out) {
    if (out) *out = nullptr;
    PTD_FAIL(PTD_ERR_UNSUPPORTED, "ptd_dn_create_strip: row-strip tiling is not built yet (DESIGN.md, multi-GPU)");
}

extern "C" ptd_status ptd_dn_forward(ptd_dn* h, const float* gbuf, float* rgb, int reset_hidden, void* stream_) {
    if (!h || !gbuf || !rgb) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_forward: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    CUDA_TRY(cudaSetDevice(h->device));
    int launches = 0;
    int nmark = 0;
    if (h->profiling) h->launch_names.clear();
    auto mark = [&](const char* name) {                                 // event after the launch just issued (and one before the first)
        if (!h->profiling) return;
        if ((int)h->events.size() <= nmark) { cudaEvent_t e; cudaEventCreate(&e); h->events.push_back(e); }
        cudaEventRecord(h->events[nmark++], st);
        if (name) h->launch_names.push_back(name);
    };
    if (reset_hidden)                                                   // forward(x, j == 0): model.py:121-128
        for (int l = 0; l < 6; ++l) CUDA_TRY(cudaMemsetAsync(h->hidden[l], 0, h->hidden_bytes[l], st));
    {
        const size_t n = (size_t)h->Hp * h->Wp;
        mark(nullptr);
        pack_gbuffer<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gbuf, h->H, h->W, h->Hp, h->Wp, h->d_in16, h->flags == PTD_DN_TF32);
        ++launches;
        mark("pack_gbuffer");
    }
    size_t pool_i = 0;
    for (size_t li = 0; li < h->layers.size(); ++li) {
        DnLayer& L = h->layers[li];
        bool pooled_in_epilogue = false;
        if (h->flags == PTD_DN_TF32) {
            ptd_status rc = tc_conv_launch(L.tc, st, &launches, &pooled_in_epilogue);
            if (rc != PTD_OK) return rc;
        } else {
            FpConvArgs a;
            a.src0 = L.src0; a.src1 = L.src1; a.c0p = L.c0p; a.c1p = L.c1p; a.upsample = L.spec.kind == DN_DEC1;
            a.H = L.H; a.W = L.W; a.w = L.d_w9; a.scale = L.d_scale; a.shift = L.d_shift; a.bias = L.d_bias;
            a.coutp = L.coutp; a.order = L.spec.lrelu_first ? 1 : 0; a.out = L.out;
            dim3 grid((L.W + FC_TW - 1) / FC_TW, (L.H + FC_TH - 1) / FC_TH, (L.coutp + FC_BN - 1) / FC_BN);
            conv3x3_fp32<<<grid, 128, 0, st>>>(a);
            ++launches;
        }
        mark(L.spec.name.c_str());
        if (L.spec.kind == DN_L2B && L.spec.level < 5) {
            const ptd_dn::Pool& p = h->pools[pool_i++];
            if (!pooled_in_epilogue) {
                const size_t n = (size_t)p.Ho * p.Wo * (p.cp / 4);
                maxpool2_nhwc<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.in, p.out, p.Ho, p.Wo, p.cp);
                ++launches;
                mark("maxpool2");
            }
        }
    }
    {
        const size_t n = (size_t)h->H * h->W;
        unpack_rgb<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->d_final, h->H, h->W, h->Wp, rgb);
        ++launches;
        mark("unpack_rgb");
    }
    h->launches = launches;
    if (h->profiling) h->timed_launches = nmark - 1;
    CUDA_TRY(cudaGetLastError());
    return PTD_OK;
}

extern "C" ptd_status ptd_dn_forward_host(ptd_dn* h, const float* gbuf_host, float* rgb_host, int reset_hidden) {
    if (!h || !gbuf_host || !rgb_host) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_forward_host: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t P = (size_t)h->H * h->W;
    CUDA_TRY(cudaMemcpy(h->d_gbuf, gbuf_host, P * 40, cudaMemcpyHostToDevice));          // main.cpp:104-105
    ptd_status rc = ptd_dn_forward(h, h->d_gbuf, h->d_rgb, reset_hidden, nullptr);
    if (rc != PTD_OK) return rc;
    CUDA_TRY(cudaMemcpy(rgb_host, h->d_rgb, P * 12, cudaMemcpyDeviceToHost));            // main.cpp:91 (.to(kCPU))
    return PTD_OK;
}

extern "C" ptd_status ptd_dn_padded_size(const ptd_dn* h, int* Hp, int* Wp) {
    if (!h) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_padded_size: null handle");
    if (Hp) *Hp = h->Hp;
    if (Wp) *Wp = h->Wp;
    return PTD_OK;
}

extern "C" ptd_status ptd_dn_dump_hidden(ptd_dn* h, int level, float* host, size_t cap, int* C, int* H, int* W) {
    if (!h || !host || level < 0 || level > 5) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_dump_hidden: bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const int c = h->hidden_c[level], hh = h->Hp >> level, ww = h->Wp >> level;
    const size_t n = (size_t)c * hh * ww;
    if (cap < n) PTD_FAIL(PTD_ERR_ARG, "ptd_dn_dump_hidden: capacity %zu < %zu", cap, n);
    float* tmp = nullptr;
    CUDA_TRY(cudaMalloc((void**)&tmp, n * 4));
    nhwc_to_nchw<<<(unsigned)((n + 255) / 256), 256>>>(h->hidden[level], c, cpad(c), hh, ww, tmp);
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
