// ptd_cli - headless drop-in for the reference executable's frame loop (Inference/src/main.cpp:47-87, :120-168).
//
//   ptd_cli SCENEFILE.txt [--weights model.ptdw] [--frames N] [--dphi RAD] [--mode tf32|f16|2xf16|fp32|3xtf32|traced] [--sort-material]
//           [--res W H] [--depth D] [--out PREFIX] [--device K] [--host-roundtrip] [--serial] [--reset-every N] [--quiet]
//
// `ptd_cli SCENEFILE.txt` is the reference's command line (main.cpp:50-56).  Every frame repeats runCuda(): the orbit camera is
// rebuilt from (zoom, phi, theta) (main.cpp:122-140), a fresh 1-spp iteration is traced (iteration == 1 every frame because
// camchanged is set again at :164) and, unless no weights were given (the reference's GROUND_TRUTH / !DENOISE_RENDER switch,
// main.cpp:40-42), the G-buffer goes through the recurrent denoiser with the hidden state carried frame to frame.  There is
// no window: the mouse drag of main.cpp:193-223 is replaced by a constant phi step per frame (--dphi) and
// cv::imshow (main.cpp:89-100) by optional PNG / PFM dumps (--out).  Host C++ only; all GPU work goes through include/ptd.h.
// Default data flow: everything stays on the device and the path trace of frame k + 1 runs on its own stream, overlapping the
// denoiser of frame k (two G-buffers, two events per frame; --serial puts both on one stream, --host-roundtrip reproduces the
// reference's PCIe round trip through host_tensor).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "ptd.h"

static int fail(const char* what) {
    fprintf(stderr, "ptd_cli: %s: %s\n", what, ptd_last_error());
    return 2;
}

// ---- minimal PNG writer (stored deflate blocks; replaces image::savePNG, image.cpp:22-39) ----------------------------
static uint32_t crc_table[256];
static void crc_init() {
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        crc_table[n] = c;
    }
}
static uint32_t crc_update(uint32_t c, const uint8_t* p, size_t n) {
    for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c;
}
static void put32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
static void chunk(FILE* f, const char* tag, const std::vector<uint8_t>& data) {
    std::vector<uint8_t> head;
    put32(head, (uint32_t)data.size());
    fwrite(head.data(), 1, 4, f);
    fwrite(tag, 1, 4, f);
    if (!data.empty()) fwrite(data.data(), 1, data.size(), f);
    uint32_t c = crc_update(0xffffffffu, (const uint8_t*)tag, 4);
    c = crc_update(c, data.data(), data.size()) ^ 0xffffffffu;
    std::vector<uint8_t> tail;
    put32(tail, c);
    fwrite(tail.data(), 1, 4, f);
}
// rgb: planar float [3][H][W]; clamp(x, 0, 1) * 255 like sendImageToPBO (pathtrace.cu:67-75)
static bool write_png(const std::string& path, const float* rgb, int W, int H) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    crc_init();
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    fwrite(sig, 1, 8, f);
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)W); put32(ihdr, (uint32_t)H);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(f, "IHDR", ihdr);
    std::vector<uint8_t> raw((size_t)H * (1 + 3 * (size_t)W));
    const size_t P = (size_t)W * H;
    for (int y = 0; y < H; ++y) {
        uint8_t* row = raw.data() + (size_t)y * (1 + 3 * (size_t)W);
        row[0] = 0;
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < 3; ++c) {
                float v = rgb[c * P + (size_t)y * W + x] * 255.0f;
                int q = (int)v;
                row[1 + 3 * x + c] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
            }
    }
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t off = 0; off < raw.size(); off += 65535) {
        const size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xff)); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xff)); z.push_back((uint8_t)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = 0; i < n; ++i) { a = (a + raw[off + i]) % 65521u; b = (b + a) % 65521u; }
    }
    put32(z, (b << 16) | a);
    chunk(f, "IDAT", z);
    chunk(f, "IEND", std::vector<uint8_t>());
    fclose(f);
    return true;
}
// planar float [3][H][W] -> PFM (bottom-up rows, little endian)
static bool write_pfm(const std::string& path, const float* rgb, int W, int H) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "PF\n%d %d\n-1.0\n", W, H);
    const size_t P = (size_t)W * H;
    std::vector<float> row(3 * (size_t)W);
    for (int y = H - 1; y >= 0; --y) {
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < 3; ++c) row[3 * x + c] = rgb[c * P + (size_t)y * W + x];
        fwrite(row.data(), 4, row.size(), f);
    }
    fclose(f);
    return true;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        printf("Usage: %s SCENEFILE.txt [--weights model.ptdw] [--frames N] [--dphi RAD] [--mode tf32|f16|2xf16|fp32|3xtf32|traced] [--sort-material]\n"
               "       [--res W H] [--depth D] [--out PREFIX] [--device K] [--host-roundtrip] [--serial] [--reset-every N] [--quiet]\n", argv[0]);   // main.cpp:50-53
        return 1;
    }
    const char* scene_file = argv[1];
    std::string weights, out_prefix, mode = "tf32";
    int frames = 1, device = 0, res_w = 0, res_h = 0, depth = 0, reset_every = 0;
    float dphi = 0.002f;
    bool sort_material = false, host_roundtrip = false, quiet = false, serial = false;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](int n) { if (i + n >= argc) { fprintf(stderr, "ptd_cli: %s needs %d value(s)\n", a.c_str(), n); exit(1); } };
        if (a == "--weights") { need(1); weights = argv[++i]; }
        else if (a == "--frames") { need(1); frames = atoi(argv[++i]); }
        else if (a == "--dphi") { need(1); dphi = (float)atof(argv[++i]); }
        else if (a == "--mode") { need(1); mode = argv[++i]; }
        else if (a == "--sort-material") sort_material = true;
        else if (a == "--res") { need(2); res_w = atoi(argv[++i]); res_h = atoi(argv[++i]); }
        else if (a == "--depth") { need(1); depth = atoi(argv[++i]); }
        else if (a == "--out") { need(1); out_prefix = argv[++i]; }
        else if (a == "--device") { need(1); device = atoi(argv[++i]); }
        else if (a == "--host-roundtrip") host_roundtrip = true;
        else if (a == "--serial") serial = true;
        else if (a == "--reset-every") { need(1); reset_every = atoi(argv[++i]); }
        else if (a == "--quiet") quiet = true;
        else { fprintf(stderr, "ptd_cli: unknown option %s\n", a.c_str()); return 1; }
    }
    unsigned dn_flags;
    if (mode == "tf32") dn_flags = PTD_DN_TF32;
    else if (mode == "fp32") dn_flags = PTD_DN_FP32;
    else if (mode == "3xtf32") dn_flags = PTD_DN_3XTF32;
    else if (mode == "f16") dn_flags = PTD_DN_F16;
    else if (mode == "2xf16") dn_flags = PTD_DN_2XF16;
    else if (mode == "traced") dn_flags = PTD_DN_FP32_BATCH_STATS;     // what the reference's TorchScript export runs; combine with --reset-every 1
    else { fprintf(stderr, "ptd_cli: unknown --mode %s\n", mode.c_str()); return 1; }

    ptd_scene* scene = nullptr;
    if (ptd_scene_load(scene_file, &scene) != PTD_OK) return fail("scene");                  // new Scene(sceneFile), main.cpp:58
    if (res_w > 0 && res_h > 0 && ptd_scene_set_resolution(scene, res_w, res_h) != PTD_OK) return fail("--res");
    if (depth > 0 && ptd_scene_set_depth(scene, depth) != PTD_OK) return fail("--depth");
    int counts[5];
    ptd_scene_counts(scene, counts);
    ptd_camera* cam = ptd_scene_camera(scene);
    const int W = cam->res_x, H = cam->res_y;
    const size_t P = (size_t)W * H;
    float zoom, phi, theta;
    ptd_camera_orbit_params(cam, &zoom, &phi, &theta);                                         // main.cpp:66-78
    if (!quiet) printf("scene %s: %d geoms, %d materials, %d faces, %dx%d, depth %d\n", scene_file, counts[0], counts[1], counts[2], W, H, counts[3]);

    if (ptd_device_count() <= device) { fprintf(stderr, "ptd_cli: CUDA device %d not available (there is no CPU fallback)\n", device); return 2; }
    ptd_pt* pt = nullptr;
    if (ptd_pt_create(scene, device, sort_material ? PTD_PT_SORT_MATERIAL : 0u, &pt) != PTD_OK) return fail("ptd_pt_create");
    ptd_dn* dn = nullptr;
    if (!weights.empty() && ptd_dn_create(weights.c_str(), H, W, device, dn_flags, &dn) != PTD_OK) return fail("ptd_dn_create");
    if (!dn && !quiet) printf("no --weights: path trace only (the reference's DENOISE_RENDER false, main.cpp:42)\n");

    cudaSetDevice(device);
    const bool pipelined = !serial && !host_roundtrip && dn != nullptr;
    float *d_gbuf[2] = {nullptr, nullptr}, *d_rgb = nullptr;
    cudaStream_t stream, s_pt;                         // `stream`: denoiser (and everything, when serial); s_pt: path trace when pipelined
    cudaEvent_t ev_pt[2], ev_dn[2];
    bool dn_recorded[2] = {false, false};
    if (cudaStreamCreate(&stream) != cudaSuccess || cudaStreamCreate(&s_pt) != cudaSuccess || cudaMalloc((void**)&d_gbuf[0], 40 * P) != cudaSuccess ||
        cudaMalloc((void**)&d_gbuf[1], 40 * P) != cudaSuccess || cudaMalloc((void**)&d_rgb, 12 * P) != cudaSuccess) {
        fprintf(stderr, "ptd_cli: device allocation failed\n");
        return 2;
    }
    for (int i = 0; i < 2; ++i) { cudaEventCreateWithFlags(&ev_pt[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_dn[i], cudaEventDisableTiming); }
    std::vector<float> h_gbuf(host_roundtrip || !out_prefix.empty() ? 10 * P : 0), h_rgb(3 * P);
    const auto t0 = std::chrono::steady_clock::now();
    for (int frame = 0; frame < frames; ++frame) {
        ptd_camera c = *cam;
        ptd_camera_orbit(&c, zoom, phi + dphi * (float)frame, theta);                          // runCuda(): main.cpp:122-140
        const int reset = frame == 0 || (reset_every > 0 && frame % reset_every == 0);        // forward(x, j == 0)
        const int gi = pipelined ? (frame & 1) : 0;
        if (host_roundtrip) {
            // the reference's exact data flow: G-buffer D2H (pathtrace.cu:525), H2D again + output D2H (main.cpp:104-105,91)
            if (ptd_pt_render_host(pt, &c, 1, h_gbuf.data()) != PTD_OK) return fail("ptd_pt_render_host");
            if (dn && ptd_dn_forward_host(dn, h_gbuf.data(), h_rgb.data(), reset) != PTD_OK) return fail("ptd_dn_forward_host");
        } else if (pipelined) {
            if (dn_recorded[gi]) cudaStreamWaitEvent(s_pt, ev_dn[gi], 0);                      // frame - 2's denoiser is done with this G-buffer
            if (ptd_pt_render(pt, &c, 1, d_gbuf[gi], s_pt) != PTD_OK) return fail("ptd_pt_render");
            cudaEventRecord(ev_pt[gi], s_pt);
            cudaStreamWaitEvent(stream, ev_pt[gi], 0);
            if (ptd_dn_forward(dn, d_gbuf[gi], d_rgb, reset, stream) != PTD_OK) return fail("ptd_dn_forward");
            cudaEventRecord(ev_dn[gi], stream);
            dn_recorded[gi] = true;
        } else {
            if (ptd_pt_render(pt, &c, 1, d_gbuf[0], stream) != PTD_OK) return fail("ptd_pt_render");
            if (dn && ptd_dn_forward(dn, d_gbuf[0], d_rgb, reset, stream) != PTD_OK) return fail("ptd_dn_forward");
        }
        if (!out_prefix.empty()) {
            if (!host_roundtrip) {
                cudaStreamSynchronize(stream);
                cudaMemcpy(h_gbuf.data(), d_gbuf[gi], 40 * P, cudaMemcpyDeviceToHost);
                if (dn) cudaMemcpy(h_rgb.data(), d_rgb, 12 * P, cudaMemcpyDeviceToHost);
            }
            char name[64];
            snprintf(name, sizeof name, "_%04d", frame);
            write_png(out_prefix + name + "_1spp.png", h_gbuf.data(), W, H);
            if (dn) { write_png(out_prefix + name + "_denoised.png", h_rgb.data(), W, H); write_pfm(out_prefix + name + "_denoised.pfm", h_rgb.data(), W, H); }
        }
    }
    cudaStreamSynchronize(s_pt);
    if (cudaStreamSynchronize(stream) != cudaSuccess) { fprintf(stderr, "ptd_cli: CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!quiet) {
        std::vector<int> live(counts[3]);
        int run = 0;
        ptd_pt_live_counts(pt, live.data(), counts[3], &run);
        printf("%d frame(s) in %.3f s (%.1f frames/s%s); live paths per bounce of the last frame:", frames, sec, frames / sec, out_prefix.empty() ? "" : ", incl. image dumps");
        for (int b = 0; b < run; ++b) printf(" %d", live[b]);
        printf("\n");
    }
    cudaFree(d_gbuf[0]); cudaFree(d_gbuf[1]); cudaFree(d_rgb); cudaStreamDestroy(stream); cudaStreamDestroy(s_pt);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(ev_pt[i]); cudaEventDestroy(ev_dn[i]); }
    ptd_dn_destroy(dn);
    ptd_pt_destroy(pt);
    ptd_scene_free(scene);
    return 0;
}
