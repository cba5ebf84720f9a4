// Internal declarations shared by the host (C++) and device (CUDA) translation units of libptd.so.
#pragma once
#include <string>
#include <vector>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "ptd.h"

static_assert(sizeof(ptd_path_segment) == 44 && sizeof(ptd_intersection) == 36 && sizeof(ptd_geom) == 248 &&
              sizeof(ptd_face) == 76 && sizeof(ptd_material) == 44 && sizeof(ptd_camera) == 84 && sizeof(ptd_aabb) == 24,
              "record layouts must match Inference/src/sceneStructs.h");

void ptd_set_error(const char* fmt, ...);
#define PTD_FAIL(code, ...) do { ptd_set_error(__VA_ARGS__); return (code); } while (0)

#ifdef __CUDACC__
// Watchdog of the device-side waits on another GPU's stores (row-strip mode: halo flags of the convs, live-count mail of pt_shade).
// A peer that died or a mis-ordered launch would otherwise spin for ever and take the GPU with it; after PTD_SPIN_TIMEOUT_NS the
// waiting kernel traps instead, the context reports a launch failure on the next CUDA call and the C ABI returns PTD_ERR_CUDA.
// tick() is called once per poll; the timer is read every 1024 polls (a poll is an L2 round trip, so about once per millisecond).
#ifndef PTD_SPIN_TIMEOUT_NS
#define PTD_SPIN_TIMEOUT_NS 20000000000ull              /* 20 s: far above any start-up skew between the ranks of one box */
#endif
struct PtdSpinGuard {
    unsigned long long t0 = 0ull; unsigned polls = 0u;
    __device__ __forceinline__ void tick() {
        if ((++polls & 1023u) != 0u) return;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0ull) t0 = t;
        else if (t - t0 > PTD_SPIN_TIMEOUT_NS) __trap();
    }
};
#endif

struct ptd_scene {
    std::vector<ptd_geom> geoms;
    std::vector<ptd_material> materials;
    std::vector<ptd_face> faces;
    ptd_aabb mesh_box;
    ptd_camera camera;
    float fovy_deg = 45.f;
    int trace_depth = 8;
    int iterations = 1;
    std::string image_name;
};

// ---- BVH over the mesh faces (new work: the reference brute-forces every face, pathtrace.cu:258-269) ----
// 32-byte node.  Interior: count == 0, left child = first, right child = first + 1.
// Leaf: count > 0, triangles [first, first + count) of the leaf-ordered triangle array.
struct PtdBvhNode {
    float bmin[3]; int first;
    float bmax[3]; int count;
};
// Leaf-ordered triangle record for traversal: vertex positions only (48 B); w of v0 holds the original face
// index (the reference's array order decides ties, pathtrace.cu:261), w of v1 the material id.
struct PtdBvhTri {
    float v0[3]; int face;
    float v1[3]; int material;
    float v2[3]; int pad;
};
// Traversal layout (64 B per INTERIOR node, the classic while-while kernel layout): both children's boxes in one fetch.
//   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   n3 = (child0, child1, -, -) as ints
// child >= 0: index of an interior node; child < 0: leaf, ~child = (first << 4) | (count - 1) into the leaf-ordered triangles.
struct PtdBvhWide { float f[16]; };
// 4-wide traversal layout (128 B per interior node), collapsed from the binary tree: a ray's traversal is a chain of DEPENDENT
// node loads (it is latency bound, not bandwidth bound), and four children per step halve the length of that chain.
//   f[0..3] lo.x of children 0..3, f[4..7] hi.x, f[8..11] lo.y, f[12..15] hi.y, f[16..19] lo.z, f[20..23] hi.z,
//   f[24..27] child codes as ints (same coding as above; an unused slot has an inverted box and is never hit), f[28..31] pad
struct PtdBvh4 { float f[32]; };
struct PtdBvh {
    std::vector<PtdBvhNode> nodes;
    std::vector<PtdBvhWide> wide;
    std::vector<PtdBvh4> wide4;
    int max_depth4 = 0;
    std::vector<PtdBvhTri> tris;
    int leaves = 0, max_leaf = 0, max_depth = 0;
};
void ptd_build_bvh(const std::vector<ptd_face>& faces, PtdBvh& out);
// 8-wide BVH with 8-bit quantised child boxes (ptd_bvh8.cpp): 80-byte nodes, five 16-byte loads per visit
struct PtdBvh8Node {
    float origin[3]; uint8_t e[3]; uint8_t imask;
    int child_base; int tri_base;            // tri_base: low 24 bits = first triangle, high 8 bits = valid-slot mask
    uint8_t meta[8];
    uint8_t qlo[3][8], qhi[3][8];
};
static_assert(sizeof(PtdBvh8Node) == 80, "BVH8q node is 80 bytes");
struct PtdBvh8 { std::vector<PtdBvh8Node> nodes; std::vector<PtdBvhTri> tris; int max_depth = 0, max_node_tris = 0; bool ok = false; };
void ptd_build_bvh8(const std::vector<ptd_face>& faces, const PtdBvh& binary, PtdBvh8& out);
void ptd_bvh8_node_hits(const PtdBvh8Node& n, const float o[3], const float idir[3], int oct, float tlim, unsigned* interior_hits, unsigned* tri_hits);
// Conservative (padded) world-space boxes of the cube / sphere geoms: a ray that misses one cannot hit the geom.
void ptd_geom_bounds(const std::vector<ptd_geom>& geoms, std::vector<ptd_aabb>& out);

// what ptd_frame_host (ptd_pt.cu) needs to know about a denoiser handle (ptd_dn.cu)
void ptd_dn_describe(const ptd_dn* h, int* device, int* H, int* W, int* strip);
// ptd_frame_submit / ptd_frame_wait (ptd_pt.cu) drive the denoiser through these: the frames in flight own the handle's state
ptd_status ptd_dn_forward_frame(ptd_dn* h, const float* gbuf, float* rgb, int reset_hidden, void* stream);
void ptd_dn_mark_inflight(ptd_dn* h, int delta);
void ptd_dn_set_sm_limit(ptd_dn* h, int sms);          /* > 0: conv grids of the following forwards use at most this many CTAs (SM partition) */

// camera helpers shared with the CLI
void ptd_camera_derive(ptd_camera& cam, float fovy_deg);
