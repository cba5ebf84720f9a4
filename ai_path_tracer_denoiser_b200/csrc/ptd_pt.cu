// HP-1: the 1-spp path-trace iteration, B200-native.
//
// Replaces pathtraceInit / pathtrace / pathtraceFree (Inference/src/pathtrace.cu:96-145, :422-528).
// Where the reference runs, per bounce, memset + computeIntersections + cudaDeviceSynchronize + shadeMaterial +
// thrust::partition (temp malloc, 3 kernels, D2H of the count) and after the loop finalGather + copy_data + a
// 40*P-byte D2H, this file runs TWO kernels per bounce and nothing after the loop:
//
//   pt_trace<FIRST>:   [ray generation at bounce 0 | ray of PathSegment idx] -> nearest hit over the geoms (shared-memory copy)
//                      and the mesh (BVH, tie-break exact) -> 36-byte ShadeableIntersection at slot idx (+ normal/depth planes
//                      at bounce 0).  Persistent warps, dynamic ray fetch, no block barrier: traversal time per ray varies 10x.
//   pt_shade<FIRST>:   coalesced smem-staged load of the PathSegment + ShadeableIntersection tile
//                      -> shade / scatter (same RNG stream: seed = hash(iter, compacted index, remainingBounces))
//                      -> G-buffer planes written in place (albedo at bounce 0, radiance at termination)
//                      -> stable stream compaction: warp-ballot ranks + block scan + decoupled look-back across tiles,
//                         survivors staged in shared memory and stored coalesced; the live count stays on the device.
//
// The PathSegment array keeps the reference's 44-byte AoS layout in HBM (so parity dumps are plain copies and the
// algorithmic bytes are the survey's 44 B read + 44 B written per live path per bounce); coalescing comes from the
// shared-memory staging, not from a layout change.  No host synchronisation happens inside a frame.
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ptd_internal.h"
#include "pt_math.cuh"

#ifndef PT_BLOCK
#define PT_BLOCK 512
#endif
#define PT_WORDS 11                 // sizeof(PathSegment) / 4
#define PT_STACK 96
#define PT_MAX_RANKS 8

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ptd_set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return PTD_ERR_CUDA; } } while (0)

struct PtKernelParams {
    const ptd_geom* geoms; const ptd_aabb* geom_bounds; int ngeoms; int geoms_in_smem;
    const ptd_material* materials; int nmaterials;
    const ptd_face* faces; int nfaces;
    const float4* nodes; const float4* tris; int use_bvh;
    ptd_aabb mesh_box;
    ptd_camera cam;
    int W, P, iter, bounce, trace_depth;   // P = pixels of THIS handle (the whole frame, or its row strip)
    int Pfull, pix0;                       // pixels of the whole frame (G-buffer plane stride); global index of this handle's first pixel
    int rank, nranks;                      // row-strip mode: position among the strips of the frame (0, 1 when not tiled)
    const unsigned long long* mail;        // [trace_depth + 1][PT_MAX_RANKS]: (epoch << 32 | live count) written by the strips above us
    unsigned long long* peer_mail[PT_MAX_RANKS];   // the other strips' mailboxes (peer memory), null for ranks <= ours
    unsigned epoch;
    const ptd_path_segment* src; ptd_path_segment* dst; ptd_path_segment* dead;
    int* counts;                    // counts[b] = live paths entering bounce b
    unsigned long long* status;     // decoupled look-back tile states of this bounce
    int* ticket;                    // ray counter of this bounce's pt_trace (dynamic fetch)
    int* ticket2;                   // dynamic tile id of this bounce's pt_shade
    ptd_intersection* isx;          // ShadeableIntersection[n] exchanged between the two kernels
    float* gbuf; float* image;
    int* sort_keys;
    ptd_path_segment* trace_paths;
    // PTD_PT_RAY_SORT (appended: the offsets of everything above are what the default kernels were validated with)
    const int* order;               // order[k] = slot of the k-th ray in spatial-bin order (pt_trace<false, true> only)
    int refill;                     // refill threshold of the binned trace kernel
    // pt_shade<.., .., true>: the survivors' bin keys and the next bounce's histogram are produced while the rays are still in registers
    ptd_aabb bin_box; int bin_bits; unsigned* bin_keys; int* bin_hist_next;
    int sstack_off;                 // pt_trace<.., .., true>: byte offset of the shared-memory traversal stacks inside the dynamic shared memory
};

using namespace ptm;

// ---- mesh traversal -------------------------------------------------------------------------------------------
// Nearest face hit with the reference's semantics: candidates are accepted in ascending face order with a strict
// `t_min > t` (pathtrace.cu:259-268), i.e. minimum t and, among equal t, the lowest face index; a geom hit with the
// same t (best_face < 0) is never displaced.
__device__ __forceinline__ void consider_face(float t, int face, float bx, float by, float& t_min, int& best_face, float& bbx, float& bby, bool& hit_face) {
    if (t > 0.0f && (t_min > t || (t_min == t && hit_face && face < best_face))) {
        t_min = t; best_face = face; bbx = bx; bby = by; hit_face = true;
    }
}

// conservative slab test against a geom's padded world box: false only when the ray certainly misses the geom.
// (idx, idy, idz) = the ray's clamped inverse direction, computed once per ray.
__device__ __forceinline__ bool ray_may_hit_box(const Ray& ray, float idx, float idy, float idz, const ptd_aabb& b) {
    const float x0 = (b.lb.x - ray.o.x) * idx, x1 = (b.ub.x - ray.o.x) * idx;
    const float y0 = (b.lb.y - ray.o.y) * idy, y1 = (b.ub.y - ray.o.y) * idy;
    const float z0 = (b.lb.z - ray.o.z) * idz, z1 = (b.ub.z - ray.o.z) * idz;
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    return !(tn * 0.9999f > tf * 1.0001f);          // NaN compares false -> "may hit"
}

// generateRayFromCamera, pathtrace.cu:155-182 (stale remainingBounces := 0 in the seed, decision D4)
__device__ __forceinline__ Ray camera_ray(const ptd_camera& cam, int W, int iter, int idx) {
    const int x = idx % W, y = idx / W;
    Rng rng = make_rng(iter, idx, 0);
    Ray ray;
    ray.o = cam.position;
    float jx = rng_uniform(rng, -0.5f, 0.5f);
    float jy = rng_uniform(rng, -0.5f, 0.5f);
    v3 a = muls(muls(cam.right, cam.pixelLength_x), fadd(ffma((float)cam.res_x, -0.5f, (float)x), jx));
    v3 b = muls(muls(cam.up, cam.pixelLength_y), fadd(ffma((float)cam.res_y, -0.5f, (float)y), jy));
    ray.d = normalize(sub(sub(cam.view, a), b));
    return ray;
}

// ---- kernel 1 of a bounce: nearest intersection (computeIntersections, pathtrace.cu:200-306) -----------------------------
// Persistent warps with dynamic ray fetch (Aila & Laine's "persistent speculative while-while"): a warp pulls rays from a
// global counter, lanes whose ray has finished are refilled as soon as fewer than TR_REFILL lanes are still traversing, and
// nothing in the kernel is a block barrier - BVH traversal lengths vary by 10x between neighbouring rays, so a block-wide
// barrier (or a fixed ray-per-thread mapping) leaves most of the machine waiting for the slowest ray of each tile.
// All loops are warp-uniform (full-mask votes decide the trip count, idle lanes are predicated off) so the warp stays
// converged.  The result is the reference's 36-byte ShadeableIntersection record at slot `idx`, i.e. this kernel and
// pt_shade exchange exactly the data computeIntersections and shadeMaterial exchange.
#define TR_BLOCK 128
#ifndef TR_REFILL
#define TR_REFILL 22
#endif
#ifndef TR_MIN_BLOCKS
#define TR_MIN_BLOCKS 8
#endif
#define PT_SENTINEL 0x76543210
#define FRAME_SLOTS 3                                   // frames ptd_frame_submit keeps in flight (G-buffer / image slots, mailboxes)

struct TraceOut {
    ptd_intersection* isx; float* gbuf; int P, W, pix0; bool write_gbuf;   // P: pixels of the whole frame
};
__device__ __forceinline__ void write_hit(const TraceOut& o, int idx, float t, v3 normal, int mat, bool outside, v3 ip) {
    ptd_intersection r;
    r.t = t; r.surfaceNormal = normalize(normal); r.materialId = mat; r.is_inside = !outside; r.pad[0] = r.pad[1] = r.pad[2] = 0; r.intersect = ip;
    o.isx[idx] = r;
    if (o.write_gbuf) {                                            // :295-304, x-mirrored like copy_data; bounce 0: pixelIndex == pix0 + idx
        const int col = (o.pix0 + idx) % o.W, row = (o.pix0 + idx) / o.W;
        const size_t m = (size_t)(o.W - col - 1) + (size_t)row * o.W;
        o.gbuf[(size_t)o.P * 3 + m] = normal.x; o.gbuf[(size_t)o.P * 4 + m] = normal.y; o.gbuf[(size_t)o.P * 5 + m] = normal.z;
        o.gbuf[(size_t)o.P * 6 + m] = t;
    }
}
__device__ __forceinline__ void write_miss(const TraceOut& o, int idx) {
    ptd_intersection r;
    memset(&r, 0, sizeof r);                                       // the reference memsets the array every bounce (:478) ...
    r.t = -1.0f;                                                   // ... and a miss only sets t (:283)
    o.isx[idx] = r;
    if (o.write_gbuf) {                                            // planes stay at the init memset's 0 (:119)
        const int col = (o.pix0 + idx) % o.W, row = (o.pix0 + idx) / o.W;
        const size_t m = (size_t)(o.W - col - 1) + (size_t)row * o.W;
        o.gbuf[(size_t)o.P * 3 + m] = 0.f; o.gbuf[(size_t)o.P * 4 + m] = 0.f; o.gbuf[(size_t)o.P * 5 + m] = 0.f; o.gbuf[(size_t)o.P * 6 + m] = 0.f;
    }
}

// SSTACK (PTD_PT_SMEM_STACK=1, opt-in): the first TR_SSTACK entries of every lane's traversal stack live in shared memory, laid out
// [entry][thread] so that a warp's push / pop is one conflict-free shared-memory access whatever the lanes' stack depths are.  In local
// memory the same push / pop touches one 128-byte line per distinct depth in the warp (~7 L1 wavefronts) - about 15 % of pt_trace's L1
// traffic.  The host probe sees at most 14 live entries on C3; deeper entries fall back to the local array.
#define TR_SSTACK 16
#define TR_PUSH(v) do { ++sp; if (SSTACK && sp < TR_SSTACK) s_stack[sp * TR_BLOCK + tid] = (v); else stack[sp] = (v); } while (0)
#define TR_POP() ((SSTACK && sp < TR_SSTACK) ? s_stack[(sp--) * TR_BLOCK + tid] : stack[sp--])
template <bool FIRST, bool BINNED = false, bool SSTACK = false>
__global__ void __launch_bounds__(TR_BLOCK, TR_MIN_BLOCKS) pt_trace(const PtKernelParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ptd_geom* s_geoms = reinterpret_cast<ptd_geom*>(smem_raw);
    int* s_stack = SSTACK ? reinterpret_cast<int*>(smem_raw + p.sstack_off) : nullptr;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = FIRST ? p.P : p.counts[p.bounce];
    if (p.geoms_in_smem) {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(p.geoms);
        uint32_t* d = reinterpret_cast<uint32_t*>(s_geoms);
        for (int i = tid; i < p.ngeoms * 62; i += TR_BLOCK) d[i] = __ldg(&g[i]);
        const uint32_t* gb = reinterpret_cast<const uint32_t*>(p.geom_bounds);
        uint32_t* db = reinterpret_cast<uint32_t*>(s_geoms + p.ngeoms);
        for (int i = tid; i < p.ngeoms * 6; i += TR_BLOCK) db[i] = __ldg(&gb[i]);
        __syncthreads();
    }
    const ptd_geom* geoms = p.geoms_in_smem ? s_geoms : p.geoms;
    const ptd_aabb* gbounds = p.geoms_in_smem ? reinterpret_cast<const ptd_aabb*>(s_geoms + p.ngeoms) : p.geom_bounds;
    TraceOut out;
    out.isx = p.isx; out.gbuf = p.gbuf; out.P = p.Pfull; out.W = p.W; out.pix0 = p.pix0; out.write_gbuf = FIRST && p.iter == 1;

    // per-lane ray state
    int idx = 0;
    bool have = false, exhausted = false, pool_empty = false;
    Ray ray; ray.o = ray.d = V(0, 0, 0);
    float idirx = 0.f, idiry = 0.f, idirz = 0.f, oodx = 0.f, oody = 0.f, oodz = 0.f;
    float t_min = FLT_MAX, bbx = 0.f, bby = 0.f;
    int best_face = -1;
    bool hit_face = false, geom_hit = false, outside = true;
    int node = PT_SENTINEL, leaf = 0, sp = 0, tri = 0, tri_end = 0;
    int nearx = 0, neary = 0, nearz = 0;
    int stack[PT_STACK];

    for (;;) {
        // ---- refill idle lanes from the global ray counter --------------------------------------------------------
        const bool need = !have && !exhausted;
        const unsigned m = __ballot_sync(FULL, need);
        if (m) {
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(p.ticket, __popc(m));
            base = __shfl_sync(FULL, base, leader);
            if (need) {
                idx = base + __popc(m & ((1u << lane) - 1u));
                if (idx >= n) {
                    exhausted = true;
                } else {
                    if (BINNED) idx = __ldg(&p.order[idx]);              // consecutive tickets = rays of one spatial / direction bin
                    have = true;
                    if (FIRST) {
                        ray = camera_ray(p.cam, p.W, p.iter, p.pix0 + idx);
                    } else {
                        const float* w = reinterpret_cast<const float*>(p.src) + (size_t)idx * PT_WORDS;
                        ray.o = V(w[0], w[1], w[2]); ray.d = V(w[3], w[4], w[5]);
                    }
                    t_min = FLT_MAX; best_face = -1; hit_face = false; outside = true; bbx = bby = 0.f;
                    v3 ip = V(0, 0, 0), normal = V(0, 0, 0), tip = V(0, 0, 0), tn = V(0, 0, 0);
                    int materialid = -1;
                    const float ooeps = 1e-30f;                                           // clamped inverse direction, shared with the BVH slabs
                    const float rix = 1.0f / (fabsf(ray.d.x) > ooeps ? ray.d.x : copysignf(ooeps, ray.d.x));
                    const float riy = 1.0f / (fabsf(ray.d.y) > ooeps ? ray.d.y : copysignf(ooeps, ray.d.y));
                    const float riz = 1.0f / (fabsf(ray.d.z) > ooeps ? ray.d.z : copysignf(ooeps, ray.d.z));
                    for (int g = 0; g < p.ngeoms; ++g) {
                        const ptd_geom* ge = &geoms[g];
                        if (!ray_may_hit_box(ray, rix, riy, riz, gbounds[g])) continue;   // a certain miss leaves t, outside untouched in the reference too
                        float t = 0.f;
                        if (ge->type == PTD_CUBE) t = boxIntersectionTest(ge, ray, tip, tn, outside);
                        else if (ge->type == PTD_SPHERE) t = sphereIntersectionTest(ge, ray, tip, tn, outside);
                        else continue;
                        if (t > 0.0f && t_min > t) { t_min = t; materialid = ge->materialid; ip = tip; normal = tn; }
                    }
                    geom_hit = materialid != -1;
                    if (geom_hit) write_hit(out, idx, t_min, normal, materialid, outside, ip);   // provisional: a nearer face overwrites it
                    node = PT_SENTINEL; leaf = 0;
                    if (p.nfaces && RayAABBintersect(ray, p.mesh_box)) {               // RAY_CULLING true, :258
                        if (p.use_bvh) {
                            idirx = rix; idiry = riy; idirz = riz;
                            oodx = ray.o.x * idirx; oody = ray.o.y * idiry; oodz = ray.o.z * idirz;
                            nearx = idirx < 0.0f; neary = idiry < 0.0f; nearz = idirz < 0.0f;     // 0: the lo plane is entered first, 1: the hi plane
                            if (SSTACK) s_stack[tid] = PT_SENTINEL; else stack[0] = PT_SENTINEL;
                            sp = 0; node = 0;
                        } else {                                                       // PTD_PT_NO_BVH: the reference's loop, test aid
                            for (int f = 0; f < p.nfaces; ++f) {
                                const ptd_face* fc = &p.faces[f];
                                float bx, by;
                                float t = triangleParam(fc->v[0], fc->v[1], fc->v[2], ray, bx, by);
                                consider_face(t, f, bx, by, t_min, best_face, bbx, bby, hit_face);
                            }
                        }
                    }
                }
            }
            pool_empty = __any_sync(FULL, exhausted);
        }
        if (!__any_sync(FULL, have)) break;

        // ---- traverse until too few lanes are busy ----------------------------------------------------------------
        for (;;) {
            // interior nodes: every lane with a node steps; the phase ends when no lane is still looking for its first leaf
            bool searching = true;
            for (;;) {
                const bool can_step = node >= 0 && node != PT_SENTINEL;
                if (!__any_sync(FULL, can_step && searching)) break;
                if (can_step) {
                    // 4-wide node (128 B): seven 16-byte read-only loads issued together.  (256-bit LDG.E.ENL2.256 loads were measured
                    // slower than 128-bit ones; 32-byte 16-bit-quantised nodes were slower too: the chain of dependent loads, not
                    // L1 request throughput, bounds this kernel - hence four children per step.)
                    const float4* np = p.nodes + 8 * (size_t)node;
                    const float4 nx = __ldg(np + nearx), fx = __ldg(np + (nearx ^ 1)), ny = __ldg(np + 2 + neary), fy = __ldg(np + 2 + (neary ^ 1));
                    const float4 nz = __ldg(np + 4 + nearz), fz = __ldg(np + 4 + (nearz ^ 1)), cc = __ldg(np + 6);
                    // slabs with the near / far plane picked per ray by the direction signs.  Conservative by construction: every box
                    // is padded by 3e-5 x the scene extent (ptd_build_bvh), i.e. by >= 3e-5 * S * |1/d| in ray-parameter units, while
                    // the rounding of fma(plane, 1/d, -o/d) is <= 1.2e-7 * |o| * |1/d| with |o| <= S inside the scene - a 250x margin.
                    // The best-t bound comes from the (differently rounded) triangle test, hence its own 1e-5 relative slack.
                    const float tlim = t_min * 1.00001f;
                    float d0, d1, d2, d3;
#define PT_SLAB(k, d)                                                                                                        \
                    {                                                                                                        \
                        const float tn = fmaxf(fmaxf(nx.k * idirx - oodx, ny.k * idiry - oody), fmaxf(nz.k * idirz - oodz, 0.0f)); \
                        const float tf = fminf(fminf(fx.k * idirx - oodx, fy.k * idiry - oody), fz.k * idirz - oodz);        \
                        d = (tn <= tf && tn <= tlim) ? tn : FLT_MAX;                                                         \
                    }
                    PT_SLAB(x, d0) PT_SLAB(y, d1) PT_SLAB(z, d2) PT_SLAB(w, d3)
#undef PT_SLAB
                    int c0 = __float_as_int(cc.x), c1 = __float_as_int(cc.y), c2 = __float_as_int(cc.z), c3 = __float_as_int(cc.w);
                    // sort the four (entry distance, child) pairs ascending: 5-comparator network
#define PT_CSWAP(da, ca, db, cb) { const bool sw = db < da; const float td = sw ? db : da; db = sw ? da : db; da = td; const int tcn = sw ? cb : ca; cb = sw ? ca : cb; ca = tcn; }
                    PT_CSWAP(d0, c0, d1, c1) PT_CSWAP(d2, c2, d3, c3) PT_CSWAP(d0, c0, d2, c2) PT_CSWAP(d1, c1, d3, c3) PT_CSWAP(d1, c1, d2, c2)
#undef PT_CSWAP
                    // the nearest hit child is visited next, the others wait on the stack, farthest at the bottom
                    if (d3 < FLT_MAX) TR_PUSH(c3);
                    if (d2 < FLT_MAX) TR_PUSH(c2);
                    if (d1 < FLT_MAX) TR_PUSH(c1);
                    if (d0 < FLT_MAX) node = c0; else node = TR_POP();
                    if (node < 0 && leaf >= 0) {            // first leaf found: postpone it and keep walking (speculatively)
                        searching = false;
                        leaf = node;
                        node = TR_POP();
                    }
                }
            }
            // leaves: one triangle per lane per trip, so lanes with short leaves do not wait for lanes with long ones
            if (leaf < 0) { const int code = ~leaf; tri = code >> 4; tri_end = tri + (code & 15) + 1; }
            for (;;) {
                const bool busy = leaf < 0;
                if (!__any_sync(FULL, busy)) break;
                if (busy) {
                    const float4* tp = p.tris + 3 * (size_t)tri;
                    const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                    float bx, by;
                    const float t = triangleParam(V(a.x, a.y, a.z), V(b.x, b.y, b.z), V(c.x, c.y, c.z), ray, bx, by);
                    consider_face(t, __float_as_int(a.w), bx, by, t_min, best_face, bbx, bby, hit_face);
                    if (++tri == tri_end) {
                        leaf = node;                        // a second leaf may be waiting in `node`
                        if (node < 0) {
                            node = TR_POP();
                            const int code = ~leaf; tri = code >> 4; tri_end = tri + (code & 15) + 1;
                        }
                    }
                }
            }
            const unsigned act = __ballot_sync(FULL, node != PT_SENTINEL);
            if (act == 0u) break;
            if (!pool_empty && __popc(act) < (BINNED ? p.refill : TR_REFILL)) break;
        }

        // ---- retire the rays that finished ------------------------------------------------------------------------
        if (have && node == PT_SENTINEL) {
            if (hit_face) {
                const ptd_face* fc = &p.faces[best_face];
                v3 ip, normal;
                triangleFinish(fc, bbx, bby, ip, normal);
                write_hit(out, idx, t_min, normal, fc->materialid, outside, ip);
            } else if (!geom_hit) {
                write_miss(out, idx);
            }
            have = false;
        }
    }
}

// bin key of a ray for PTD_PT_RAY_SORT (see "coherent scheduling of the secondary rays" below): Morton cell of the origin | direction octant
__device__ __forceinline__ unsigned ray_bin(const float* w, const ptd_aabb& box, int bits) {
    const float cells = (float)(1 << bits);
    const float fx = (w[0] - box.lb.x) / fmaxf(box.ub.x - box.lb.x, 1e-20f), fy = (w[1] - box.lb.y) / fmaxf(box.ub.y - box.lb.y, 1e-20f);
    const float fz = (w[2] - box.lb.z) / fmaxf(box.ub.z - box.lb.z, 1e-20f);
    const int mx = (1 << bits) - 1;
    const unsigned cx = (unsigned)min(max((int)(fx * cells), 0), mx), cy = (unsigned)min(max((int)(fy * cells), 0), mx), cz = (unsigned)min(max((int)(fz * cells), 0), mx);
    unsigned morton = 0;
    for (int b = 0; b < bits; ++b) morton |= (((cx >> b) & 1u) << (3 * b)) | (((cy >> b) & 1u) << (3 * b + 1)) | (((cz >> b) & 1u) << (3 * b + 2));
    const unsigned octant = (w[3] < 0.f ? 1u : 0u) | (w[4] < 0.f ? 2u : 0u) | (w[5] < 0.f ? 4u : 0u);
    return (morton << 3) | octant;                                    // cell-major: concurrently running warps work in the same region
}

// ---- block-wide helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- -DPTD_SHADE_PROF: per-phase globaltimer sums of the shade kernels (tools/shade_prof.py), compiled out of the product ----------------
#ifdef PTD_SHADE_PROF
__device__ unsigned long long g_shade_prof[16];
__device__ __forceinline__ unsigned long long prof_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define SH_PROF_BEGIN() unsigned long long prof_t = prof_now()
#define SH_PROF(k) do { if (threadIdx.x == 0) { const unsigned long long t_ = prof_now(); atomicAdd(&g_shade_prof[k], t_ - prof_t); prof_t = t_; } } while (0)
#define SH_PROF_TILE() do { if (threadIdx.x == 0) atomicAdd(&g_shade_prof[15], 1ull); } while (0)
extern "C" int ptd_debug_shade_prof(unsigned long long* out16, int reset) {
    if (out16) cudaMemcpyFromSymbol(out16, g_shade_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(g_shade_prof, z, sizeof z); }
    return 0;
}
#else
#define SH_PROF_BEGIN() do {} while (0)
#define SH_PROF(k) do {} while (0)
#define SH_PROF_TILE() do {} while (0)
#endif

// ---- shadeMaterial (pathtrace.cu:333-390) + finalGather (:393-402) + copy_data (:81-94) for ONE path ---------------------------------
// `seed_index` is the path's index in the frame-wide compacted array (what the reference seeds its RNG with, :351).  Returns whether the
// path lives on; a path that ends here adds its throughput to the accumulation image and emits the radiance planes - every segment
// terminates exactly once per iteration, so the two extra passes over P the reference makes are not needed.
template <bool FIRST>
__device__ __forceinline__ bool shade_path(const PtKernelParams& p, int seed_index, float isx_t, v3 isx_n, int isx_mat, v3 ip,
                                           Ray& ray, v3& color, int pixelIndex, int& rb) {
    const bool hit = isx_t >= 0;
    const int col = pixelIndex % p.W, row = pixelIndex / p.W;
    const size_t mirrored = (size_t)(p.W - col - 1) + (size_t)row * p.W;              // x-mirror of copy_data / :297-299
    if (isx_t > 0.0f) {
        Rng rng = make_rng(p.iter, seed_index, rb);
        const ptd_material m = p.materials[isx_mat];
        if (m.emittance > 0.0f) {
            rb = 0;
            color = muls(mulv(color, m.color), m.emittance);
        } else {
            scatterRay(ray, color, ip, isx_n, m, rng);
            --rb;
        }
    } else {
        color = V(0, 0, 0);
        rb = 0;
    }
    if (FIRST && p.iter == 1) {                                                         // :379-387
        p.gbuf[(size_t)p.Pfull * 7 + mirrored] = hit ? color.x : 0.f;
        p.gbuf[(size_t)p.Pfull * 8 + mirrored] = hit ? color.y : 0.f;
        p.gbuf[(size_t)p.Pfull * 9 + mirrored] = hit ? color.z : 0.f;
    }
    if (rb > 0) return true;
    float* img = p.image + (size_t)(pixelIndex - p.pix0) * 3;
    v3 acc = add(p.iter != 1 ? V(img[0], img[1], img[2]) : V(0.f, 0.f, 0.f), color);
    img[0] = acc.x; img[1] = acc.y; img[2] = acc.z;
    const float fi = (float)p.iter;
    p.gbuf[mirrored] = __fdiv_rn(acc.x, fi);
    p.gbuf[(size_t)p.Pfull + mirrored] = __fdiv_rn(acc.y, fi);
    p.gbuf[(size_t)p.Pfull * 2 + mirrored] = __fdiv_rn(acc.z, fi);
    return false;
}

// ---- kernel 2 of a bounce, one tile per block (PTD_PT_SHADE_TILED=1; round 1's kernel, kept for A/B): shadeMaterial (:333-390) + thrust::partition (:505) + finalGather / copy_data for the paths that end here
// WIDE (PTD_PT_WIDE_LOOKBACK=1, opt-in): the decoupled look-back reads PT_BLOCK predecessor states per trip with the whole block
// instead of 32 with one warp.  With ~300 tiles in flight the one-warp walk is ~10 dependent L2 round trips per tile while the other
// 15 warps sit at the barrier (ncu: stall_barrier dominates pt_shade); one block-wide trip covers every tile in flight.
// KEYS (PTD_PT_RAY_SORT, next bounce binned): every survivor's bin key is written at its compacted index and counted into the next
// bounce's histogram here, where the new ray is still in registers - the separate ray_bin_hist pass (a 40 MB re-read) is not needed.
template <bool FIRST, bool WIDE = false, bool KEYS = false>
__global__ void __launch_bounds__(PT_BLOCK) pt_shade_tiled(const PtKernelParams p) {
    __shared__ __align__(16) uint32_t s_words[PT_BLOCK * PT_WORDS];                  // 5632 B staging (loads, then compacted stores)
    __shared__ __align__(16) uint32_t s_isx[PT_BLOCK * 9];                           // 4608 B: the tile's ShadeableIntersections
    __shared__ int s_tile, s_warp_kept[PT_BLOCK / 32], s_excl, s_goff;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SH_PROF_BEGIN();
    const int n = FIRST ? p.P : p.counts[p.bounce];
    if (tid == 0) s_tile = atomicAdd(p.ticket2, 1);                      // dynamic tile id: look-back predecessors are always scheduled
    __syncthreads();
    const int tile = s_tile;
    const int base = tile * PT_BLOCK;
    if (base >= n) {
        if (n == 0 && tile == 0 && tid == 0) {                            // nothing alive here: still tell the strips below
            const unsigned long long m = (unsigned long long)p.epoch << 32;
            for (int r = p.rank + 1; r < p.nranks; ++r)
                if (p.peer_mail[r]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_mail[r] + (size_t)(p.bounce + 1) * PT_MAX_RANKS + p.rank), "l"(m) : "memory");
        }
        return;
    }
    if (tid == 0) {
        // Row-strip mode: the reference seeds its RNG with the index in the frame-wide compacted array (pathtrace.cu:351).  Strips
        // are contiguous pixel ranges and compaction is stable, so that index is (live paths of the strips above) + local index;
        // the strips above published their live counts of this bounce into our mailbox (peer stores) when they compacted.
        int goff = FIRST ? p.pix0 : 0;
        if (!FIRST) {
            PtdSpinGuard guard;                                            // traps after PTD_SPIN_TIMEOUT_NS instead of hanging the GPU
            for (int r = 0; r < p.rank; ++r) {
                unsigned long long m;
                for (;;) {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(m) : "l"(p.mail + (size_t)p.bounce * PT_MAX_RANKS + r) : "memory");
                    if ((unsigned)(m >> 32) == p.epoch) break;
                    guard.tick();
                }
                goff += (int)(unsigned)m;
            }
        }
        s_goff = goff;                                                    // read after the next __syncthreads
    }
    SH_PROF(0);                                                            // ticket + mail
    const int valid = min(PT_BLOCK, n - base);
    const int idx = base + tid;
    const bool active = tid < valid;

    // ---- 1. this tile's path segments and intersections (coalesced, staged through shared memory) ----------------
    Ray ray; v3 color; int pixelIndex = 0, rb = 0;
    ray.o = ray.d = color = V(0, 0, 0);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.isx) + (size_t)base * 9;
        if (valid == PT_BLOCK && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {   // 4608 B tile
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(s_isx);
            for (int i = tid; i < PT_BLOCK * 9 / 4; i += PT_BLOCK) d4[i] = s4[i];
        } else {
            for (int i = tid; i < valid * 9; i += PT_BLOCK) s_isx[i] = src[i];
        }
    }
    if (FIRST) {
        if (active) {
            ray = camera_ray(p.cam, p.W, p.iter, p.pix0 + idx);
            color = V(1.0f, 1.0f, 1.0f);
            pixelIndex = p.pix0 + idx;
            rb = p.trace_depth;
        }
        __syncthreads();
    } else {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.src) + (size_t)base * PT_WORDS;
        if (valid == PT_BLOCK) {                                           // 5632 B tile, 16-B aligned: 352 uint4
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(s_words);
            for (int i = tid; i < PT_BLOCK * PT_WORDS / 4; i += PT_BLOCK) d4[i] = s4[i];
        } else {
            for (int i = tid; i < valid * PT_WORDS; i += PT_BLOCK) s_words[i] = src[i];
        }
        __syncthreads();
        if (active) {
            const float* w = reinterpret_cast<const float*>(s_words) + tid * PT_WORDS;   // stride 11 words: conflict free
            ray.o = V(w[0], w[1], w[2]); ray.d = V(w[3], w[4], w[5]); color = V(w[6], w[7], w[8]);
            pixelIndex = __float_as_int(w[9]); rb = __float_as_int(w[10]);
        }
    }
    float isx_t = -1.0f; v3 isx_n = V(0, 0, 0), ip = V(0, 0, 0); int isx_mat = 0;
    if (active) {
        const float* w = reinterpret_cast<const float*>(s_isx) + tid * 9;                // stride 9 words: conflict free
        isx_t = w[0]; isx_n = V(w[1], w[2], w[3]); isx_mat = __float_as_int(w[4]); ip = V(w[6], w[7], w[8]);
    }
    __syncthreads();                                                       // staging buffer is reused for the stores below
    SH_PROF(1);                                                            // loads
    if (p.trace_paths && active) {
        ptd_path_segment ps;
        ps.ray.origin = ray.o; ps.ray.direction = ray.d; ps.color = color; ps.pixelIndex = pixelIndex; ps.remainingBounces = rb;
        p.trace_paths[(size_t)p.bounce * p.P + idx] = ps;
    }

    bool keep = false;
    if (active) {
        if (p.sort_keys) p.sort_keys[idx] = isx_mat;                                   // key of the UN-compacted slot (:509 quirk)
        keep = shade_path<FIRST>(p, s_goff + idx, isx_t, isx_n, isx_mat, ip, ray, color, pixelIndex, rb);
    }

    SH_PROF(2);                                                            // shade
    // ---- 3. stable stream compaction (thrust::partition, :505) -----------------------------------------------
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int lane_rank = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp_kept[warp] = __popc(ballot);
    __syncthreads();
    SH_PROF(3);                                                            // the block's slowest warp
    int warp_off = 0, block_kept = 0;
#pragma unroll
    for (int w = 0; w < PT_BLOCK / 32; ++w) { if (w < warp) warp_off += s_warp_kept[w]; block_kept += s_warp_kept[w]; }
    if (WIDE) {
        __shared__ int s_first[PT_BLOCK / 32], s_pending[PT_BLOCK / 32], s_sum[PT_BLOCK / 32];
        int excl = 0;
        if (tile == 0) {
            if (tid == 0) st_status(&p.status[0], (2ull << 32) | (unsigned)block_kept);
        } else {
            if (tid == 0) st_status(&p.status[tile], (1ull << 32) | (unsigned)block_kept);
            int j = tile - 1;                                              // thread `tid` looks at predecessor j - tid
            for (;;) {
                const int t = j - tid;
                const unsigned long long sv = t >= 0 ? ld_status(&p.status[t]) : (2ull << 32);   // before tile 0: prefix 0
                const unsigned st = (unsigned)(sv >> 32);
                const unsigned has_prefix = __ballot_sync(0xffffffffu, st == 2u);
                const unsigned not_ready = __ballot_sync(0xffffffffu, st == 0u);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;                    // nearest prefix inside this warp's 32 predecessors
                const unsigned needed = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);    // lanes 0 .. first (all of them when there is none)
                int v = ((needed >> lane) & 1u) ? (int)(unsigned)sv : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) { s_first[warp] = first; s_pending[warp] = (not_ready & needed) != 0u; s_sum[warp] = v; }
                __syncthreads();
                int total = 0; bool pending = false, found = false;
#pragma unroll 1
                for (int w = 0; w < PT_BLOCK / 32; ++w) {                  // warps in look-back order, up to the first one that saw a prefix
                    pending |= s_pending[w] != 0; total += s_sum[w];
                    if (s_first[w] < 32) { found = true; break; }
                }
                __syncthreads();                                           // the three arrays are rewritten by the next trip
                if (pending) continue;                                     // block-uniform: a needed predecessor has not published yet
                excl += total;
                if (found) break;
                j -= PT_BLOCK;
            }
            if (tid == 0) st_status(&p.status[tile], (2ull << 32) | (unsigned)(excl + block_kept));
        }
        if (tid == 0) {
            s_excl = excl;
            if (base + PT_BLOCK >= n) {                                                       // last tile publishes the live count and mails it
                p.counts[p.bounce + 1] = excl + block_kept;
                const unsigned long long m = ((unsigned long long)p.epoch << 32) | (unsigned)(excl + block_kept);
                for (int r = p.rank + 1; r < p.nranks; ++r)
                    if (p.peer_mail[r]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_mail[r] + (size_t)(p.bounce + 1) * PT_MAX_RANKS + p.rank), "l"(m) : "memory");
            }
        }
    } else if (warp == 0) {
        // decoupled look-back, one WARP wide: state 1 = tile aggregate, 2 = inclusive prefix (value in the low 32 bits).  Each trip
        // reads the 32 predecessors' states at once; with ~1000 tiles in flight a one-thread walk was a chain of hundreds of
        // dependent L2 reads and left the rest of the block at the barrier (ncu: stall_barrier 19.6 of 33 warps per issue).
        int excl = 0;
        if (tile == 0) {
            if (lane == 0) st_status(&p.status[0], (2ull << 32) | (unsigned)block_kept);
        } else {
            if (lane == 0) st_status(&p.status[tile], (1ull << 32) | (unsigned)block_kept);
            int j = tile - 1;                                              // nearest predecessor of this window
            for (;;) {
                const int t = j - lane;
                const unsigned long long sv = t >= 0 ? ld_status(&p.status[t]) : (2ull << 32);   // before tile 0: prefix 0
                const unsigned st = (unsigned)(sv >> 32);
                const unsigned has_prefix = __ballot_sync(0xffffffffu, st == 2u);
                const unsigned not_ready = __ballot_sync(0xffffffffu, st == 0u);
                const int first_prefix = has_prefix ? __ffs(has_prefix) - 1 : 31;
                const unsigned needed = first_prefix == 31 ? 0xffffffffu : ((2u << first_prefix) - 1u);   // lanes 0 .. first_prefix
                if (not_ready & needed) continue;                           // a predecessor has not published yet: look again
                int v = lane <= first_prefix ? (int)(unsigned)sv : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                excl += v;
                if (has_prefix) break;
                j -= 32;
            }
            if (lane == 0) st_status(&p.status[tile], (2ull << 32) | (unsigned)(excl + block_kept));
        }
        if (lane == 0) {
            s_excl = excl;
            if (base + PT_BLOCK >= n) {                                                       // last tile publishes the live count
                p.counts[p.bounce + 1] = excl + block_kept;
                const unsigned long long m = ((unsigned long long)p.epoch << 32) | (unsigned)(excl + block_kept);
                for (int r = p.rank + 1; r < p.nranks; ++r)                                    // ... and mails it to the strips below (NVLink peer stores)
                    if (p.peer_mail[r]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_mail[r] + (size_t)(p.bounce + 1) * PT_MAX_RANKS + p.rank), "l"(m) : "memory");
            }
        }
    }
    const int local_rank = warp_off + lane_rank;
    if (keep) {
        float* w = reinterpret_cast<float*>(s_words) + local_rank * PT_WORDS;
        w[0] = ray.o.x; w[1] = ray.o.y; w[2] = ray.o.z; w[3] = ray.d.x; w[4] = ray.d.y; w[5] = ray.d.z;
        w[6] = color.x; w[7] = color.y; w[8] = color.z; w[9] = __int_as_float(pixelIndex); w[10] = __int_as_float(rb);
    }
    __syncthreads();
    SH_PROF(4);                                                            // look-back
    const int excl = s_excl;
    {
        uint32_t* dst = reinterpret_cast<uint32_t*>(p.dst) + (size_t)excl * PT_WORDS;
        const int nwords = block_kept * PT_WORDS;
        for (int i = tid; i < nwords; i += PT_BLOCK) dst[i] = s_words[i];               // contiguous, 128 B per warp store
    }
    if (KEYS && keep) {
        const float w6[6] = {ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z};
        const unsigned k = ray_bin(w6, p.bin_box, p.bin_bits);
        p.bin_keys[excl + local_rank] = k;
        atomicAdd(&p.bin_hist_next[k], 1);
    }
    if (p.dead && active && !keep) {
        // rejected items end up behind the survivors in REVERSE order (thrust CUDA back end) and are never moved again
        const int rej_before = base - excl + (tid - local_rank);
        ptd_path_segment ps;
        ps.ray.origin = ray.o; ps.ray.direction = ray.d; ps.color = color; ps.pixelIndex = pixelIndex; ps.remainingBounces = rb;
        p.dead[(size_t)(n - 1 - rej_before)] = ps;
    }
    SH_PROF(5);                                                            // stores issued
    SH_PROF_TILE();
}

// ---- kernel 2 of a bounce, pipelined (the default) ---------------------------------------------------------------------------------
// Same work, same results, same status / count / mail protocol as pt_shade_tiled, organised as a software pipeline.  Per-tile phase times
// of the tiled kernel (globaltimer stamps, C3, profiles/r4b_shade_phases.json): 3.6 us waiting for the tile's loads, 1.6 us of shading, 4.2 us
// in the decoupled look-back, 0.6 us of stores - the block does one of these at a time.  Here a persistent block (2 per SM) keeps three
// tiles in flight:
//   * tile k + 1 is fetched by two bulk copies (cp.async.bulk, one mbarrier per buffer) issued before tile k is touched;
//   * tile k is shaded by 16 warps, its survivors are staged compacted in shared memory and its aggregate is published at once;
//   * the look-back of tile k runs on a 17th warp while the 16 shade tile k + 1; tile k's survivors are stored when its exclusive prefix
//     arrives, one iteration later.
// Buffers: 2 x ShadeableIntersection tile (18 KB), 3 x PathSegment tile (22 KB: being filled / being shaded and staged / waiting for its
// prefix) = 102 KB per block.  Tile ids are still taken from the bounce's ticket in the order blocks ask for them and every block works
// through its tiles in increasing order, publishing a tile's aggregate before it waits for anything, so the lowest unpublished tile is
// always being shaded by a resident block: the look-back cannot deadlock.
// When the frame-wide index of a dropped path is needed at once (PT_KEEP_TERMINATED's reject array, the ray-sort keys) the block waits for
// the prefix of the tile it has just shaded instead (`defer` = false); everything else is the same code.
constexpr int SH_THREADS = PT_BLOCK + 32;
constexpr int SH_ISX_WORDS = PT_BLOCK * 9, SH_PATH_WORDS = PT_BLOCK * PT_WORDS;
constexpr int SH_SMEM_BYTES = (3 * SH_PATH_WORDS + 2 * SH_ISX_WORDS) * 4;

__device__ __forceinline__ uint32_t sh_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sh_mbar_init(uint64_t* b, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sh_smem(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void sh_mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sh_smem(b)) : "memory"); }
__device__ __forceinline__ void sh_mbar_expect(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sh_smem(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void sh_mbar_wait(uint64_t* b, unsigned parity) {
    PtdSpinGuard guard;                                                    // traps instead of hanging the GPU
    for (;;) {
        unsigned ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(sh_smem(b)), "r"(parity) : "memory");
        if (ok) return;
        guard.tick();
    }
}
__device__ __forceinline__ void sh_bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sh_smem(dst)), "l"(src), "r"(bytes), "r"(sh_smem(bar)) : "memory");
}
__device__ __forceinline__ void sh_bar_shaders() { asm volatile("bar.sync 1, %0;" ::"n"(PT_BLOCK) : "memory"); }   // the 16 shading warps only

template <bool FIRST, bool KEYS = false>
__global__ void __launch_bounds__(SH_THREADS, 1024 / PT_BLOCK) pt_shade(const PtKernelParams p) {
    extern __shared__ __align__(128) uint32_t sh_dyn[];
    uint32_t* const s_path_buf = sh_dyn;                                   // [3][SH_PATH_WORDS]
    uint32_t* const s_isx_buf = sh_dyn + 3 * SH_PATH_WORDS;                // [2][SH_ISX_WORDS]
    __shared__ __align__(8) uint64_t bar_full[2], bar_req[2], bar_done[2];
    __shared__ int s_next, s_goff, s_warp_kept[PT_BLOCK / 32], s_req_tile[2], s_req_agg[2], s_excl[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = FIRST ? p.P : p.counts[p.bounce];
    const bool defer = !(KEYS || p.dead != nullptr);
    const uint32_t* const g_isx = reinterpret_cast<const uint32_t*>(p.isx);
    const uint32_t* const g_src = reinterpret_cast<const uint32_t*>(p.src);
    // a full tile whose two source ranges are 16-byte aligned comes by bulk copy; the last (partial) tile of a bounce is loaded by the block
    const bool aligned = (reinterpret_cast<uintptr_t>(g_isx) & 15) == 0 && (FIRST || (reinterpret_cast<uintptr_t>(g_src) & 15) == 0);
    auto bulk_tile = [&](int tile) { return aligned && (tile + 1) * PT_BLOCK <= n; };
    auto fetch = [&](int tile, int slot2, int slot3) {                     // one thread
        sh_mbar_expect(&bar_full[slot2], (SH_ISX_WORDS + (FIRST ? 0 : SH_PATH_WORDS)) * 4);
        sh_bulk_load(s_isx_buf + slot2 * SH_ISX_WORDS, g_isx + (size_t)tile * SH_ISX_WORDS, SH_ISX_WORDS * 4, &bar_full[slot2]);
        if (!FIRST) sh_bulk_load(s_path_buf + slot3 * SH_PATH_WORDS, g_src + (size_t)tile * SH_PATH_WORDS, SH_PATH_WORDS * 4, &bar_full[slot2]);
    };

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { sh_mbar_init(&bar_full[i], 1); sh_mbar_init(&bar_req[i], 1); sh_mbar_init(&bar_done[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int t0 = atomicAdd(p.ticket2, 1);                            // dynamic tile ids: look-back predecessors are always scheduled
        s_next = t0;
        if (bulk_tile(t0)) fetch(t0, 0, 0);
        if (n == 0 && t0 == 0) {                                           // nothing alive here: still tell the strips below
            const unsigned long long m = (unsigned long long)p.epoch << 32;
            for (int r = p.rank + 1; r < p.nranks; ++r)
                if (p.peer_mail[r]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_mail[r] + (size_t)(p.bounce + 1) * PT_MAX_RANKS + p.rank), "l"(m) : "memory");
        }
        // Row-strip mode: the reference seeds its RNG with the index in the frame-wide compacted array (pathtrace.cu:351).  Strips are
        // contiguous pixel ranges and compaction is stable, so that index is (live paths of the strips above) + local index; the strips
        // above published their live counts of this bounce into our mailbox (peer stores) when they compacted.
        int goff = FIRST ? p.pix0 : 0;
        if (!FIRST && (t0 * PT_BLOCK < n)) {
            PtdSpinGuard guard;
            for (int r = 0; r < p.rank; ++r) {
                unsigned long long m;
                for (;;) {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(m) : "l"(p.mail + (size_t)p.bounce * PT_MAX_RANKS + r) : "memory");
                    if ((unsigned)(m >> 32) == p.epoch) break;
                    guard.tick();
                }
                goff += (int)(unsigned)m;
            }
        }
        s_goff = goff;
    }
    __syncthreads();

    if (warp == PT_BLOCK / 32) {
        // ---- the look-back warp: one request per tile of this block, in order; (tile < 0) ends it --------------------------------------
        for (int q = 0;; ++q) {
            const int slot = q & 1;
            sh_mbar_wait(&bar_req[slot], (q >> 1) & 1);
            const int tile = s_req_tile[slot], agg = s_req_agg[slot];
            if (tile < 0) break;
            // decoupled look-back, one warp wide: state 1 = tile aggregate, 2 = inclusive prefix (value in the low 32 bits); each trip reads
            // 32 predecessors' states at once
            int excl = 0;
            if (tile > 0) {
                PtdSpinGuard guard;
                int j = tile - 1;                                          // nearest predecessor of this window
                for (;;) {
                    const int t = j - lane;
                    const unsigned long long sv = t >= 0 ? ld_status(&p.status[t]) : (2ull << 32);   // before tile 0: prefix 0
                    const unsigned st = (unsigned)(sv >> 32);
                    const unsigned has_prefix = __ballot_sync(0xffffffffu, st == 2u);
                    const unsigned not_ready = __ballot_sync(0xffffffffu, st == 0u);
                    const int first_prefix = has_prefix ? __ffs(has_prefix) - 1 : 31;
                    const unsigned needed = first_prefix == 31 ? 0xffffffffu : ((2u << first_prefix) - 1u);   // lanes 0 .. first_prefix
                    if (not_ready & needed) { guard.tick(); continue; }    // a predecessor has not published yet: look again (64 states per trip, a
                                                                           // back-off, tile ids taken further ahead: all measured slower)
                    int v = lane <= first_prefix ? (int)(unsigned)sv : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    excl += v;
                    if (has_prefix) break;
                    j -= 32;
                }
                if (lane == 0) st_status(&p.status[tile], (2ull << 32) | (unsigned)(excl + agg));
            }
            if (lane == 0) {
                if ((tile + 1) * PT_BLOCK >= n) {                          // last tile publishes the live count ...
                    p.counts[p.bounce + 1] = excl + agg;
                    const unsigned long long m = ((unsigned long long)p.epoch << 32) | (unsigned)(excl + agg);
                    for (int r = p.rank + 1; r < p.nranks; ++r)            // ... and mails it to the strips below (NVLink peer stores)
                        if (p.peer_mail[r]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_mail[r] + (size_t)(p.bounce + 1) * PT_MAX_RANKS + p.rank), "l"(m) : "memory");
                }
                s_excl[slot] = excl;
                sh_mbar_arrive(&bar_done[slot]);
            }
            __syncwarp();
        }
        return;
    }

    // ---- the 16 shading warps -----------------------------------------------------------------------------------------------------------
    auto store_tile = [&](const uint32_t* words, int kept, int excl) {     // contiguous, 128 B per warp store
        uint32_t* dst = reinterpret_cast<uint32_t*>(p.dst) + (size_t)excl * PT_WORDS;
        const int nwords = kept * PT_WORDS;
        for (int i = tid; i < nwords; i += PT_BLOCK) dst[i] = words[i];
    };
    int cur = s_next, k = 0, prev_kept = 0;
    SH_PROF_BEGIN();
    for (;; ++k) {
        const int base = cur * PT_BLOCK;
        if (base >= n) break;
        const int valid = min(PT_BLOCK, n - base);
        const int idx = base + tid;
        const bool active = tid < valid;
        const int i2 = k & 1, i3 = k % 3;
        uint32_t* const s_isx = s_isx_buf + i2 * SH_ISX_WORDS;
        uint32_t* const s_words = s_path_buf + i3 * SH_PATH_WORDS;

        // ---- A. next tile id; its data starts to move now --------------------------------------------------------------------------
        if (tid == 0) {
            const int nx = atomicAdd(p.ticket2, 1);                        // (taking ids further ahead delays everybody's look-back: a held id is an unpublished predecessor)
            s_next = nx;
            if (bulk_tile(nx)) fetch(nx, i2 ^ 1, (k + 1) % 3);
        }
        SH_PROF(0);                                                        // ticket + prefetch issue (+ the previous iteration's closing barrier)
        // ---- B. this tile's path segments and intersections ------------------------------------------------------------------------
        if (bulk_tile(cur)) {
            sh_mbar_wait(&bar_full[i2], (k >> 1) & 1);
        } else {
            const uint32_t* a = g_isx + (size_t)base * 9;
            for (int i = tid; i < valid * 9; i += PT_BLOCK) s_isx[i] = a[i];
            if (!FIRST) {
                const uint32_t* b = g_src + (size_t)base * PT_WORDS;
                for (int i = tid; i < valid * PT_WORDS; i += PT_BLOCK) s_words[i] = b[i];
            }
            sh_bar_shaders();
        }
        SH_PROF(1);                                                        // wait for the tile's data
        Ray ray; v3 color; int pixelIndex = 0, rb = 0;
        ray.o = ray.d = color = V(0, 0, 0);
        float isx_t = -1.0f; v3 isx_n = V(0, 0, 0), ip = V(0, 0, 0); int isx_mat = 0;
        if (active) {
            if (FIRST) {
                ray = camera_ray(p.cam, p.W, p.iter, p.pix0 + idx);
                color = V(1.0f, 1.0f, 1.0f);
                pixelIndex = p.pix0 + idx;
                rb = p.trace_depth;
            } else {
                const float* w = reinterpret_cast<const float*>(s_words) + tid * PT_WORDS;   // stride 11 words: conflict free
                ray.o = V(w[0], w[1], w[2]); ray.d = V(w[3], w[4], w[5]); color = V(w[6], w[7], w[8]);
                pixelIndex = __float_as_int(w[9]); rb = __float_as_int(w[10]);
            }
            const float* w = reinterpret_cast<const float*>(s_isx) + tid * 9;                // stride 9 words: conflict free
            isx_t = w[0]; isx_n = V(w[1], w[2], w[3]); isx_mat = __float_as_int(w[4]); ip = V(w[6], w[7], w[8]);
        }
        sh_bar_shaders();                                                  // s_words becomes the staging buffer of this tile's survivors
        const int nxt = s_next;
        if (p.trace_paths && active) {
            ptd_path_segment ps;
            ps.ray.origin = ray.o; ps.ray.direction = ray.d; ps.color = color; ps.pixelIndex = pixelIndex; ps.remainingBounces = rb;
            p.trace_paths[(size_t)p.bounce * p.P + idx] = ps;
        }
        // ---- C. shade ------------------------------------------------------------------------------------------------------------------
        bool keep = false;
        if (active) {
            if (p.sort_keys) p.sort_keys[idx] = isx_mat;                   // key of the UN-compacted slot (:509 quirk)
            keep = shade_path<FIRST>(p, s_goff + idx, isx_t, isx_n, isx_mat, ip, ray, color, pixelIndex, rb);
        }
        SH_PROF(2);                                                        // registers, barrier, shade
        // ---- D. stable stream compaction (thrust::partition, :505): ranks inside the tile, aggregate out at once ----------------------
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        const int lane_rank = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) s_warp_kept[warp] = __popc(ballot);
        sh_bar_shaders();
        int warp_off = 0, block_kept = 0;
#pragma unroll
        for (int w = 0; w < PT_BLOCK / 32; ++w) { if (w < warp) warp_off += s_warp_kept[w]; block_kept += s_warp_kept[w]; }
        if (tid == 0) {
            st_status(&p.status[cur], ((cur == 0 ? 2ull : 1ull) << 32) | (unsigned)block_kept);
            s_req_tile[i2] = cur; s_req_agg[i2] = block_kept;
            sh_mbar_arrive(&bar_req[i2]);
        }
        const int local_rank = warp_off + lane_rank;
        if (keep) {
            float* w = reinterpret_cast<float*>(s_words) + local_rank * PT_WORDS;
            w[0] = ray.o.x; w[1] = ray.o.y; w[2] = ray.o.z; w[3] = ray.d.x; w[4] = ray.d.y; w[5] = ray.d.z;
            w[6] = color.x; w[7] = color.y; w[8] = color.z; w[9] = __int_as_float(pixelIndex); w[10] = __int_as_float(rb);
        }
        SH_PROF(3);                                                        // slowest warp, scan, publish, stage
        // ---- E. stores: the tile before this one (its prefix has had a whole iteration to arrive), or this one when it cannot wait -------
        if (defer) {
            if (k > 0) {
                sh_mbar_wait(&bar_done[i2 ^ 1], ((k - 1) >> 1) & 1);
                SH_PROF(4);                                                // wait for the previous tile's prefix
                store_tile(s_path_buf + ((k + 2) % 3) * SH_PATH_WORDS, prev_kept, s_excl[i2 ^ 1]);
            }
        } else {
            sh_bar_shaders();                                              // this tile's staging is complete
            sh_mbar_wait(&bar_done[i2], (k >> 1) & 1);
            const int excl = s_excl[i2];
            store_tile(s_words, block_kept, excl);
            if (KEYS && keep) {
                const float w6[6] = {ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z};
                const unsigned key = ray_bin(w6, p.bin_box, p.bin_bits);
                p.bin_keys[excl + local_rank] = key;
                atomicAdd(&p.bin_hist_next[key], 1);
            }
            if (p.dead && active && !keep) {
                // rejected items end up behind the survivors in REVERSE order (thrust CUDA back end) and are never moved again
                const int rej_before = base - excl + (tid - local_rank);
                ptd_path_segment ps;
                ps.ray.origin = ray.o; ps.ray.direction = ray.d; ps.color = color; ps.pixelIndex = pixelIndex; ps.remainingBounces = rb;
                p.dead[(size_t)(n - 1 - rej_before)] = ps;
            }
        }
        prev_kept = block_kept;
        SH_PROF(5);                                                        // stores issued
        SH_PROF_TILE();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // our reads / writes of the buffers before the next bulk copies into them
        sh_bar_shaders();
        cur = nxt;
    }
    if (defer && k > 0) {
        sh_mbar_wait(&bar_done[(k - 1) & 1], ((k - 1) >> 1) & 1);
        store_tile(s_path_buf + ((k - 1) % 3) * SH_PATH_WORDS, prev_kept, s_excl[(k - 1) & 1]);
    }
    if (tid == 0) { s_req_tile[k & 1] = -1; sh_mbar_arrive(&bar_req[k & 1]); }   // ends the look-back warp
}

// ---- PTD_PT_RAY_SORT: coherent scheduling of the secondary rays ---------------------------------------------------------
// After a diffuse bounce neighbouring slots of the PathSegment array hold rays that start all over the scene and point anywhere:
// the 32 lanes of a trace warp walk 32 different parts of the BVH (ncu: 16 of 32 lanes active, one L1 wavefront per lane per node
// load).  This mode bins the live rays of a bounce by (Morton cell of the origin, direction octant) with a counting sort over
// slot INDICES - histogram, exclusive scan, scatter - and the trace kernel takes its rays in bin order (p.order).  Only the order
// in which rays are TRACED changes: every ShadeableIntersection is still written to its ray's own slot, so the PathSegment arrays,
// the compaction and the RNG indices - everything the parity tests compare - are untouched.
__global__ void ray_bin_hist(const ptd_path_segment* __restrict__ paths, const int* __restrict__ count, ptd_aabb box, int bits,
                             unsigned* __restrict__ keys, int* __restrict__ hist) {
    const int n = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned k = ray_bin(reinterpret_cast<const float*>(paths) + (size_t)i * PT_WORDS, box, bits);
        keys[i] = k;
        atomicAdd(&hist[k], 1);
    }
}
__global__ void ray_bin_scatter(const unsigned* __restrict__ keys, const int* __restrict__ count, int* __restrict__ offsets, int* __restrict__ order) {
    const int n = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        order[atomicAdd(&offsets[keys[i]], 1)] = i;                     // the order inside a bin does not matter
}

// PTD_PT_GATED_MAIL: one warp, no shared memory - lane r waits for strip r's live count of this bounce.  When it exits the mail
// is in place, so the wait loop at the top of pt_shade (unchanged) passes on its first read and no 512-thread shade block with
// 41 KB of shared memory ever sits on an SM waiting for another GPU.
__global__ void __launch_bounds__(32) pt_mail_gate(const unsigned long long* mail, int bounce, int rank, unsigned epoch) {
    if ((int)threadIdx.x >= rank) return;
    PtdSpinGuard guard;
    for (;;) {
        unsigned long long m;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(m) : "l"(mail + (size_t)bounce * PT_MAX_RANKS + threadIdx.x) : "memory");
        if ((unsigned)(m >> 32) == epoch) break;
        guard.tick();
    }
}

// ---- material sort (SORT_MATERIAL, pathtrace.cu:508-510): stable counting sort of dst[0,k) by keys[0,k) --------------
#define SORT_TILE 256
__global__ void sort_hist(const int* __restrict__ keys, const int* __restrict__ count, int nbins, int nblocks, int* __restrict__ hist) {
    extern __shared__ int s_h[];
    const int k = *count;
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    const int i = blockIdx.x * SORT_TILE + threadIdx.x;
    if (i < k) atomicAdd(&s_h[min(max(keys[i], 0), nbins - 1)], 1);
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) hist[(size_t)b * nblocks + blockIdx.x] = s_h[b];
}
__global__ void sort_scan(int* __restrict__ hist, int total) {           // exclusive scan, one block of 1024 threads
    __shared__ int s_sum[1024];
    const int per = (total + 1023) / 1024;
    const int lo = min(total, (int)threadIdx.x * per), hi = min(total, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += hist[i];
    s_sum[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int v = threadIdx.x >= off ? s_sum[threadIdx.x - off] : 0;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_sum[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) { int v = hist[i]; hist[i] = run; run += v; }
}
__global__ void sort_scatter(const int* __restrict__ keys, const int* __restrict__ count, int nbins, int nblocks, const int* __restrict__ offsets,
                             const ptd_path_segment* __restrict__ in, ptd_path_segment* __restrict__ out) {
    __shared__ int s_k[SORT_TILE];
    const int k = *count;
    const int i = blockIdx.x * SORT_TILE + threadIdx.x;
    const int key = i < k ? min(max(keys[i], 0), nbins - 1) : -1;
    s_k[threadIdx.x] = key;
    __syncthreads();
    if (i >= k) return;
    int rank = 0;
    for (int j = 0; j < (int)threadIdx.x; ++j) rank += (s_k[j] == key);
    out[offsets[(size_t)key * nblocks + blockIdx.x] + rank] = in[i];
}

__global__ void export_rgba8(const float* __restrict__ image, int W, int H, int iter, uchar4* __restrict__ pbo) {   // sendImageToPBO :59-79
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;   // H: rows of this handle
    if (x >= W || y >= H) return;
    const int index = x + y * W;
    const float* pix = image + (size_t)index * 3;
    int r = min(max((int)(pix[0] / iter * 255.0f), 0), 255);
    int g = min(max((int)(pix[1] / iter * 255.0f), 0), 255);
    int b = min(max((int)(pix[2] / iter * 255.0f), 0), 255);
    pbo[index] = make_uchar4((unsigned char)r, (unsigned char)g, (unsigned char)b, 0);
}

// ---- handle -----------------------------------------------------------------------------------------------------
struct ptd_pt {
    int device = 0;
    unsigned flags = 0;
    int W = 0, H = 0, P = 0, depth = 0, ngeoms = 0, nmaterials = 0, nfaces = 0, ntiles = 0;   // P: pixels of this handle (frame or strip)
    int Pfull = 0, row0 = 0, rows = 0;                        // whole-frame pixels; this handle's image rows [row0, row0 + rows)
    int rank = 0, nranks = 1; unsigned epoch = 0;
    unsigned long long* d_mail = nullptr;                     // [FRAME_SLOTS][depth + 1][PT_MAX_RANKS], written by the strips above (peer stores)
    unsigned long long* peer_mail[PT_MAX_RANKS] = {nullptr}; bool peer_ipc[PT_MAX_RANKS] = {false};
    ptd_camera cam;
    ptd_aabb mesh_box;
    ptd_geom* d_geoms = nullptr; ptd_aabb* d_geom_bounds = nullptr; ptd_material* d_materials = nullptr; ptd_face* d_faces = nullptr;
    float4* d_nodes = nullptr; float4* d_tris = nullptr;
    ptd_path_segment* d_paths[3] = {nullptr, nullptr, nullptr};
    ptd_path_segment* d_dead = nullptr;
    ptd_intersection* d_isx = nullptr;
    int trace_blocks = 0;
    float* d_image = nullptr; float* d_gbuf_own = nullptr; float* d_frame_rgb = nullptr;
    unsigned char* d_ctl = nullptr; size_t ctl_bytes = 0;     // counts | tickets | status (memset once per frame)
    int* d_counts = nullptr; int* d_ticket = nullptr; unsigned long long* d_status = nullptr;
    int* d_keys = nullptr; int* d_hist = nullptr; int sort_blocks = 0;
    ptd_path_segment* d_trace_paths = nullptr; ptd_intersection* d_trace_isx = nullptr;
    int final_buf = 0, cur = 0, nmark = 0;
    cudaStream_t host_stream[2] = {nullptr, nullptr}; cudaEvent_t host_event = nullptr;   // ptd_pt_render_host
    int launches = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> events;
    int timed_launches = 0;
    int bvh_nodes = 0, bvh_leaves = 0, bvh_max_leaf = 0, bvh_max_depth = 0;
    // ptd_frame_submit / ptd_frame_wait: FRAME_SLOTS frame slots, three streams (path trace, denoise + frame copy, G-buffer copy)
    cudaStream_t fr_stream[3] = {nullptr, nullptr, nullptr};
    int fr_sm_pt = 0, fr_sm_dn = 0;                          // PTD_FRAME_SM_SPLIT: SMs of the path-trace / denoiser partition (0 = the whole GPU, shared)
    void* fr_green[2] = {nullptr, nullptr};                  // the two green contexts (CUgreenCtx)
    int trace_per_sm = 4, shade_per_sm = 2;
    float* fr_gbuf[FRAME_SLOTS] = {}; float* fr_rgb[FRAME_SLOTS] = {};
#ifdef PTD_FRAME_SPANS
    cudaEvent_t sp_ev[FRAME_SLOTS][4] = {};                            // debug build: PT start / end, DN start / end of the slot's frame (timing events)
#endif
    cudaEvent_t fr_ev_pt[FRAME_SLOTS] = {}, fr_ev_done[FRAME_SLOTS] = {}, fr_ev_gcopy[FRAME_SLOTS] = {}, fr_ev_rcopy[FRAME_SLOTS] = {};
    bool fr_has_gcopy[FRAME_SLOTS] = {}, fr_has_rcopy[FRAME_SLOTS] = {};
    long long fr_submitted = 0, fr_waited = 0;
    ptd_dn* fr_dn[FRAME_SLOTS] = {};                   // the denoiser handle of the frame in each slot (its in-flight count is ours to drop)
    cudaEvent_t fr_ev_t0 = nullptr, fr_ev_t1 = nullptr; bool fr_timer_armed = false; cudaStream_t fr_last_dn = nullptr;   // ptd_frame_timer
    bool wide_lookback = false;                              // PTD_PT_WIDE_LOOKBACK=1 (tiled shade kernel only)
    bool shade_tiled = false;                                // PTD_PT_SHADE_TILED=1: one tile per block (round 1's kernel) instead of the pipelined one
    int shade_blocks = 296;                                  // persistent blocks of the pipelined shade kernel (2 per SM)
    bool smem_stack = false;                                 // PTD_PT_SMEM_STACK=1
    // PTD_PT_RAY_SORT
    bool bin_fused = true;
    int bin_bits = 0, bin_refill = TR_REFILL, nbins = 0, bin_from = 2;   // bounce 1 is still origin-coherent by pixel order: binning starts at bounce 2
    ptd_aabb bin_box;
    unsigned* d_bin_keys = nullptr; int* d_bin_order = nullptr; int* d_bin_hist = nullptr;   // hist: [depth][nbins] inside d_ctl (zeroed with it)
};

extern "C" int ptd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static void frame_green_destroy(void* green_ctx);
extern "C" void ptd_pt_destroy(ptd_pt* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_geoms); cudaFree(h->d_geom_bounds); cudaFree(h->d_materials); cudaFree(h->d_faces); cudaFree(h->d_nodes); cudaFree(h->d_tris);
    for (int i = 0; i < 3; ++i) cudaFree(h->d_paths[i]);
    cudaFree(h->d_dead); cudaFree(h->d_isx); cudaFree(h->d_image); cudaFree(h->d_gbuf_own); cudaFree(h->d_ctl); cudaFree(h->d_keys); cudaFree(h->d_hist);
    cudaFree(h->d_trace_paths); cudaFree(h->d_trace_isx); cudaFree(h->d_mail); cudaFree(h->d_bin_keys); cudaFree(h->d_bin_order); cudaFree(h->d_frame_rgb);
    for (int r = 0; r < PT_MAX_RANKS; ++r) if (h->peer_mail[r] && h->peer_ipc[r]) cudaIpcCloseMemHandle(h->peer_mail[r]);
    if (h->host_stream[0]) { cudaStreamDestroy(h->host_stream[0]); cudaStreamDestroy(h->host_stream[1]); cudaEventDestroy(h->host_event); }
    for (int i = 0; i < 3; ++i) if (h->fr_stream[i]) { cudaStreamSynchronize(h->fr_stream[i]); cudaStreamDestroy(h->fr_stream[i]); }
    for (int i = 0; i < 2; ++i) if (h->fr_green[i]) frame_green_destroy(h->fr_green[i]);
    for (int i = 0; i < FRAME_SLOTS; ++i) {
        cudaFree(h->fr_gbuf[i]); cudaFree(h->fr_rgb[i]);
        if (h->fr_ev_pt[i]) cudaEventDestroy(h->fr_ev_pt[i]);
        if (h->fr_ev_done[i]) cudaEventDestroy(h->fr_ev_done[i]);
        if (h->fr_ev_rcopy[i]) cudaEventDestroy(h->fr_ev_rcopy[i]);
        if (h->fr_ev_gcopy[i]) cudaEventDestroy(h->fr_ev_gcopy[i]);
    }
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    delete h;
}

static ptd_status pt_create(const ptd_scene* sc, int device, unsigned flags, int row0, int rows, ptd_pt** out);
extern "C" ptd_status ptd_pt_create(const ptd_scene* sc, int device, unsigned flags, ptd_pt** out) {
    if (!sc || !out) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create: null argument");
    return pt_create(sc, device, flags, 0, sc->camera.res_y, out);
}
// Row-strip mode: this handle traces image rows [row0, row0 + rows) of the frame.  Pixels are independent; the
// only coupling is the frame-wide compacted index in the RNG seed, restored on the device from the live counts the strips
// mail each other (pt_shade).  Material sort permutes paths across the whole frame and is not available in strip mode.
extern "C" ptd_status ptd_pt_create_strip(const ptd_scene* sc, int device, unsigned flags, int row0, int rows, ptd_pt** out) {
    if (!sc || !out) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create_strip: null argument");
    if (row0 < 0 || rows < 1 || row0 + rows > sc->camera.res_y) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create_strip: rows [%d, %d) outside the %d-row frame", row0, row0 + rows, sc->camera.res_y);
    if (flags & PTD_PT_SORT_MATERIAL) PTD_FAIL(PTD_ERR_UNSUPPORTED, "ptd_pt_create_strip: material sort is frame-wide and not available in row-strip mode");
    return pt_create(sc, device, flags, row0, rows, out);
}
static ptd_status pt_create(const ptd_scene* sc, int device, unsigned flags, int row0, int rows, ptd_pt** out) {
    *out = nullptr;
    if (ptd_device_count() <= device || device < 0) PTD_FAIL(PTD_ERR_CUDA, "ptd_pt_create: CUDA device %d not available (no CPU fallback exists)", device);
    if (sc->trace_depth < 1 || sc->trace_depth > 1023) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create: trace depth %d out of range", sc->trace_depth);
    for (const ptd_geom& g : sc->geoms)
        if (g.materialid < 0 || g.materialid >= (int)sc->materials.size()) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create: geom material id %d out of range", g.materialid);
    for (const ptd_face& f : sc->faces)
        if (f.materialid < 0 || f.materialid >= (int)sc->materials.size()) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create: mesh material id %d out of range", f.materialid);
    CUDA_TRY(cudaSetDevice(device));
    ptd_pt* h = new ptd_pt();
    if (const char* e = getenv("PTD_PT_RAY_SORT")) { if (atoi(e) > 0) flags |= PTD_PT_RAY_SORT; }   // tuning: opt in without touching the caller
    if (!(sc->faces.size() > 0) || (flags & PTD_PT_NO_BVH)) flags &= ~(unsigned)PTD_PT_RAY_SORT;      // binning only pays for BVH traversal
    h->device = device; h->flags = flags;
    if (const char* e = getenv("PTD_PT_WIDE_LOOKBACK")) h->wide_lookback = atoi(e) > 0;
    if (const char* e = getenv("PTD_PT_SHADE_TILED")) h->shade_tiled = atoi(e) > 0;
    if (h->wide_lookback) h->shade_tiled = true;
    if (const char* e = getenv("PTD_PT_SMEM_STACK")) h->smem_stack = atoi(e) > 0;
    h->cam = sc->camera; h->mesh_box = sc->mesh_box;
    h->W = sc->camera.res_x; h->H = sc->camera.res_y; h->Pfull = h->W * h->H; h->depth = sc->trace_depth;
    h->row0 = row0; h->rows = rows; h->P = h->W * rows;
    h->ngeoms = (int)sc->geoms.size(); h->nmaterials = (int)sc->materials.size(); h->nfaces = (int)sc->faces.size();
    h->ntiles = (h->P + PT_BLOCK - 1) / PT_BLOCK;
    const size_t P = (size_t)h->P;
#define ALLOC(ptr, bytes) do { if (cudaMalloc((void**)&(ptr), (bytes) ? (bytes) : 16) != cudaSuccess) { ptd_set_error("ptd_pt_create: cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(cudaGetLastError())); ptd_pt_destroy(h); return PTD_ERR_CUDA; } } while (0)
#define UPLOAD(dst, src, bytes) do { if ((bytes) && cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice) != cudaSuccess) { ptd_set_error("ptd_pt_create: upload failed: %s", cudaGetErrorString(cudaGetLastError())); ptd_pt_destroy(h); return PTD_ERR_CUDA; } } while (0)
    ALLOC(h->d_geoms, sizeof(ptd_geom) * sc->geoms.size());
    UPLOAD(h->d_geoms, sc->geoms.data(), sizeof(ptd_geom) * sc->geoms.size());
    {
        std::vector<ptd_aabb> gb;
        ptd_geom_bounds(sc->geoms, gb);
        ALLOC(h->d_geom_bounds, sizeof(ptd_aabb) * gb.size());
        UPLOAD(h->d_geom_bounds, gb.data(), sizeof(ptd_aabb) * gb.size());
    }
    ALLOC(h->d_materials, sizeof(ptd_material) * sc->materials.size());
    UPLOAD(h->d_materials, sc->materials.data(), sizeof(ptd_material) * sc->materials.size());
    ALLOC(h->d_faces, sizeof(ptd_face) * sc->faces.size());
    UPLOAD(h->d_faces, sc->faces.data(), sizeof(ptd_face) * sc->faces.size());
    if (h->nfaces && !(flags & PTD_PT_NO_BVH)) {
        PtdBvh bvh;
        ptd_build_bvh(sc->faces, bvh);
        h->bvh_nodes = (int)bvh.nodes.size(); h->bvh_leaves = bvh.leaves; h->bvh_max_leaf = bvh.max_leaf; h->bvh_max_depth = bvh.max_depth;
        if (3 * bvh.max_depth4 + 2 > PT_STACK) { ptd_set_error("ptd_pt_create: BVH depth %d exceeds traversal stack %d", bvh.max_depth4, PT_STACK); ptd_pt_destroy(h); return PTD_ERR_UNSUPPORTED; }
        ALLOC(h->d_nodes, sizeof(PtdBvh4) * bvh.wide4.size());
        UPLOAD(h->d_nodes, bvh.wide4.data(), sizeof(PtdBvh4) * bvh.wide4.size());
        ALLOC(h->d_tris, sizeof(PtdBvhTri) * bvh.tris.size());
        UPLOAD(h->d_tris, bvh.tris.data(), sizeof(PtdBvhTri) * bvh.tris.size());
    }
    const bool sort = (flags & PTD_PT_SORT_MATERIAL) != 0;
    for (int i = 0; i < (sort ? 3 : 2); ++i) ALLOC(h->d_paths[i], sizeof(ptd_path_segment) * P);
    if (flags & PTD_PT_KEEP_TERMINATED) { ALLOC(h->d_dead, sizeof(ptd_path_segment) * P); cudaMemset(h->d_dead, 0, sizeof(ptd_path_segment) * P); }
    ALLOC(h->d_image, sizeof(float) * 3 * P);
    cudaMemset(h->d_image, 0, sizeof(float) * 3 * P);
    if (!(flags & PTD_PT_TRACE)) ALLOC(h->d_isx, sizeof(ptd_intersection) * P);
    // control block: counts[depth+1] | tickets[2*depth] | status[depth][ntiles]
    size_t off_counts = 0, off_ticket = ((size_t)(h->depth + 1) * 4 + 15) / 16 * 16, off_status = off_ticket + ((size_t)h->depth * 8 + 15) / 16 * 16;
    h->ctl_bytes = off_status + (size_t)h->depth * h->ntiles * 8;
    const size_t off_bins = h->ctl_bytes;
    if (flags & PTD_PT_RAY_SORT) {
        h->bin_bits = 4;                                                // 4096 cells x 8 octants = 32768 bins, ~28 rays per bin at 720p
        if (const char* e = getenv("PTD_PT_RAY_SORT_BITS")) { const int v = atoi(e); if (v >= 1 && v <= 5) h->bin_bits = v; }
        if (const char* e = getenv("PTD_PT_RAY_SORT_REFILL")) { const int v = atoi(e); if (v >= 1 && v <= 32) h->bin_refill = v; }
        if (const char* e = getenv("PTD_PT_RAY_SORT_UNFUSED")) h->bin_fused = !(atoi(e) > 0);
        if (const char* e = getenv("PTD_PT_RAY_SORT_FROM")) { const int v = atoi(e); if (v >= 1 && v <= 1023) h->bin_from = v; }
        h->nbins = 8 << (3 * h->bin_bits);
        h->ctl_bytes += (size_t)h->depth * h->nbins * 4;
        // cells over the box of everything a ray can start from: the mesh and the geoms
        h->bin_box = sc->mesh_box;
        std::vector<ptd_aabb> gb;
        ptd_geom_bounds(sc->geoms, gb);
        for (const ptd_aabb& b : gb) {
            if (!(b.lb.x > -1e30f && b.ub.x < 1e30f)) continue;       // a degenerate transform's "never cull" box
            h->bin_box.lb.x = std::min(h->bin_box.lb.x, b.lb.x); h->bin_box.lb.y = std::min(h->bin_box.lb.y, b.lb.y); h->bin_box.lb.z = std::min(h->bin_box.lb.z, b.lb.z);
            h->bin_box.ub.x = std::max(h->bin_box.ub.x, b.ub.x); h->bin_box.ub.y = std::max(h->bin_box.ub.y, b.ub.y); h->bin_box.ub.z = std::max(h->bin_box.ub.z, b.ub.z);
        }
        ALLOC(h->d_bin_keys, sizeof(unsigned) * P);
        ALLOC(h->d_bin_order, sizeof(int) * P);
    }
    ALLOC(h->d_ctl, h->ctl_bytes);
    // FRAME_SLOTS mailboxes, used in turn (frame number mod FRAME_SLOTS): with at most FRAME_SLOTS frames in flight per rank (ptd_frame_submit)
    // a strip above that already renders frame k + 1 or k + 2 writes another box, and cannot reach frame k + FRAME_SLOTS before every strip has
    // read frame k's counts (its denoiser of frame k, which gates the slot, depends on every other strip's denoiser of frame k and so on their
    // path trace of frame k)
    ALLOC(h->d_mail, sizeof(unsigned long long) * FRAME_SLOTS * (size_t)(h->depth + 1) * PT_MAX_RANKS);
    cudaMemset(h->d_mail, 0, sizeof(unsigned long long) * FRAME_SLOTS * (size_t)(h->depth + 1) * PT_MAX_RANKS);
    h->d_counts = (int*)(h->d_ctl + off_counts); h->d_ticket = (int*)(h->d_ctl + off_ticket); h->d_status = (unsigned long long*)(h->d_ctl + off_status);
    if (flags & PTD_PT_RAY_SORT) h->d_bin_hist = (int*)(h->d_ctl + off_bins);
    if (sort) {
        h->sort_blocks = (h->P + SORT_TILE - 1) / SORT_TILE;
        ALLOC(h->d_keys, sizeof(int) * P);
        ALLOC(h->d_hist, sizeof(int) * (size_t)std::max(h->nmaterials, 1) * h->sort_blocks);
    }
    if (flags & PTD_PT_TRACE) {
        ALLOC(h->d_trace_paths, sizeof(ptd_path_segment) * P * h->depth);
        ALLOC(h->d_trace_isx, sizeof(ptd_intersection) * P * h->depth);
    }
#undef ALLOC
#undef UPLOAD
    {
        int sms = 148, per_sm = 8;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const size_t geom_smem = (sizeof(ptd_geom) + sizeof(ptd_aabb)) * std::min(h->ngeoms, 64);
        if (h->smem_stack) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pt_trace<false, false, true>, TR_BLOCK, ((geom_smem + 15) & ~(size_t)15) + TR_SSTACK * TR_BLOCK * 4);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pt_trace<false>, TR_BLOCK, geom_smem);
        if (const char* e = getenv("PTD_TRACE_BLOCKS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }   // tuning knob: leave room for a concurrent kernel
        h->trace_blocks = std::max(1, std::min(sms * std::max(per_sm, 1), (h->P + TR_BLOCK - 1) / TR_BLOCK));   // persistent: every resident warp pulls rays
        h->trace_per_sm = std::max(per_sm, 1);
        CUDA_TRY(cudaFuncSetAttribute(pt_shade<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(pt_shade<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(pt_shade<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(pt_shade<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_SMEM_BYTES));
        int sh_per_sm = 2;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sh_per_sm, pt_shade<false>, SH_THREADS, SH_SMEM_BYTES);
        if (const char* e = getenv("PTD_SHADE_BLOCKS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < sh_per_sm) sh_per_sm = v; }
        h->shade_blocks = sms * std::max(sh_per_sm, 1);
        h->shade_per_sm = std::max(sh_per_sm, 1);
    }
    CUDA_TRY(cudaDeviceSynchronize());
    *out = h;
    return PTD_OK;
}

// Bounces [first, last) of one iteration; first == 0 also starts the frame.  ptd_pt_render runs them all; a same-process strip
// group issues them bounce by bounce over the strips (ptd_pt_render_group).
static ptd_status pt_run(ptd_pt* h, const ptd_camera* cam, int iter, float* gbuf, cudaStream_t st, int first, int last) {
    CUDA_TRY(cudaSetDevice(h->device));
    if (cam && (cam->res_x != h->W || cam->res_y != h->H)) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_render: camera resolution %dx%d differs from the handle's %dx%d", cam->res_x, cam->res_y, h->W, h->H);
    if (!gbuf) {
        if (!h->d_gbuf_own) {
            CUDA_TRY(cudaMalloc((void**)&h->d_gbuf_own, sizeof(float) * 10 * (size_t)h->Pfull));
            CUDA_TRY(cudaMemset(h->d_gbuf_own, 0, sizeof(float) * 10 * (size_t)h->Pfull));
            CUDA_TRY(cudaDeviceSynchronize());        // the memset runs on the legacy stream; `st` may be a non-blocking stream that does not wait for it
        }
        gbuf = h->d_gbuf_own;
    }
    PtKernelParams p;
    memset(&p, 0, sizeof p);
    p.geoms = h->d_geoms; p.geom_bounds = h->d_geom_bounds; p.ngeoms = h->ngeoms; p.geoms_in_smem = h->ngeoms <= 64;
    p.materials = h->d_materials; p.nmaterials = h->nmaterials;
    p.faces = h->d_faces; p.nfaces = h->nfaces;
    p.nodes = h->d_nodes; p.tris = h->d_tris; p.use_bvh = h->d_nodes != nullptr;
    p.mesh_box = h->mesh_box;
    p.cam = cam ? *cam : h->cam;
    p.W = h->W; p.P = h->P; p.iter = iter; p.trace_depth = h->depth;
    p.Pfull = h->Pfull; p.pix0 = h->row0 * h->W; p.rank = h->rank; p.nranks = h->nranks;
    p.counts = h->d_counts; p.gbuf = gbuf; p.image = h->d_image; p.dead = h->d_dead;
    p.sort_keys = h->d_keys; p.trace_paths = h->d_trace_paths;
    size_t smem = p.geoms_in_smem ? (sizeof(ptd_geom) + sizeof(ptd_aabb)) * h->ngeoms : 0;
    if (h->smem_stack) { p.sstack_off = (int)((smem + 15) & ~(size_t)15); smem = (size_t)p.sstack_off + TR_SSTACK * TR_BLOCK * 4; }
    const bool sort = (h->flags & PTD_PT_SORT_MATERIAL) != 0;
    // keys are written at compacted indices by pt_shade unless a material sort permutes the paths afterwards (or PTD_PT_RAY_SORT_UNFUSED=1)
    const bool bin_fused = !sort && h->bin_fused;
    auto mark = [&]() {
        if (!h->profiling) return;
        if ((int)h->events.size() <= h->nmark) { cudaEvent_t e; cudaEventCreate(&e); h->events.push_back(e); }
        cudaEventRecord(h->events[h->nmark++], st);
    };
    if (first == 0) {
        h->epoch += 1;
        CUDA_TRY(cudaMemsetAsync(h->d_ctl, 0, h->ctl_bytes, st));
        h->cur = 0; h->launches = 0; h->nmark = 0;
        mark();
    }
    p.epoch = h->epoch;
    const size_t mail_box = (size_t)(h->epoch % FRAME_SLOTS) * (size_t)(h->depth + 1) * PT_MAX_RANKS;          // this frame's mailbox (see pt_create)
    p.mail = h->d_mail + mail_box;
    for (int r = 0; r < PT_MAX_RANKS; ++r) p.peer_mail[r] = h->peer_mail[r] ? h->peer_mail[r] + mail_box : nullptr;
    // on the path-trace partition of a split GPU (PTD_FRAME_SM_SPLIT) the persistent grids are sized for that partition
    const bool part = h->fr_sm_pt > 0 && st == h->fr_stream[0];
    const int trace_blocks = part ? std::min(h->trace_blocks, h->fr_sm_pt * h->trace_per_sm) : h->trace_blocks;
    const int shade_blocks = part ? std::min(h->shade_blocks, h->fr_sm_pt * h->shade_per_sm) : h->shade_blocks;
    for (int b = first; b < last && b < h->depth; ++b) {
        const int cur = h->cur, nxt = (cur + 1) % (sort ? 3 : 2);
        p.bounce = b;
        p.src = h->d_paths[cur]; p.dst = h->d_paths[nxt];
        p.status = h->d_status + (size_t)b * h->ntiles;
        p.ticket = h->d_ticket + 2 * b; p.ticket2 = h->d_ticket + 2 * b + 1;
        p.isx = h->d_trace_isx ? h->d_trace_isx + (size_t)b * h->P : h->d_isx;
        if (b == 0) {
            if (h->smem_stack) pt_trace<true, false, true><<<trace_blocks, TR_BLOCK, smem, st>>>(p);
            else pt_trace<true><<<trace_blocks, TR_BLOCK, smem, st>>>(p);
        }
        else if ((h->flags & PTD_PT_RAY_SORT) && b >= h->bin_from) {
            // bin the live rays of this bounce, then trace them in bin order (timed together with the trace kernel they serve)
            int* hist = h->d_bin_hist + (size_t)b * h->nbins;
            const int blocks = std::min((h->P + 255) / 256, 148 * 8);
            if (!bin_fused) {                                           // else bounce b - 1's pt_shade wrote the keys and this histogram (b >= bin_from >= 1)
                ray_bin_hist<<<blocks, 256, 0, st>>>(p.src, h->d_counts + b, h->bin_box, h->bin_bits, h->d_bin_keys, hist);
                h->launches += 1;
            }
            sort_scan<<<1, 1024, 0, st>>>(hist, h->nbins);
            ray_bin_scatter<<<blocks, 256, 0, st>>>(h->d_bin_keys, h->d_counts + b, hist, h->d_bin_order);
            p.order = h->d_bin_order; p.refill = h->bin_refill;
            if (h->smem_stack) pt_trace<false, true, true><<<trace_blocks, TR_BLOCK, smem, st>>>(p);
            else pt_trace<false, true><<<trace_blocks, TR_BLOCK, smem, st>>>(p);
            h->launches += 2;
        }
        else if (h->smem_stack) pt_trace<false, false, true><<<trace_blocks, TR_BLOCK, smem, st>>>(p);
        else pt_trace<false><<<trace_blocks, TR_BLOCK, smem, st>>>(p);
        mark();
        if (b > 0 && h->rank > 0 && (h->flags & PTD_PT_GATED_MAIL)) {     // (timed together with the shade kernel it gates)
            pt_mail_gate<<<1, 32, 0, st>>>(p.mail, b, h->rank, h->epoch);
            h->launches += 1;
        }
        // the next bounce is binned: this pt_shade also writes its survivors' keys and the next histogram
        const bool keys = bin_fused && (h->flags & PTD_PT_RAY_SORT) && b + 1 >= h->bin_from && b + 1 < h->depth;
        if (keys) { p.bin_box = h->bin_box; p.bin_bits = h->bin_bits; p.bin_keys = h->d_bin_keys; p.bin_hist_next = h->d_bin_hist + (size_t)(b + 1) * h->nbins; }
        if (h->shade_tiled) {
            if (keys && h->wide_lookback) {
                if (b == 0) pt_shade_tiled<true, true, true><<<h->ntiles, PT_BLOCK, 0, st>>>(p);
                else pt_shade_tiled<false, true, true><<<h->ntiles, PT_BLOCK, 0, st>>>(p);
            } else if (keys) {
                if (b == 0) pt_shade_tiled<true, false, true><<<h->ntiles, PT_BLOCK, 0, st>>>(p);
                else pt_shade_tiled<false, false, true><<<h->ntiles, PT_BLOCK, 0, st>>>(p);
            } else if (h->wide_lookback) {
                if (b == 0) pt_shade_tiled<true, true><<<h->ntiles, PT_BLOCK, 0, st>>>(p);
                else pt_shade_tiled<false, true><<<h->ntiles, PT_BLOCK, 0, st>>>(p);
            }
            else if (b == 0) pt_shade_tiled<true><<<h->ntiles, PT_BLOCK, 0, st>>>(p);
            else pt_shade_tiled<false><<<h->ntiles, PT_BLOCK, 0, st>>>(p);
        } else {
            const int blocks = std::min(h->ntiles, shade_blocks);
            if (keys) {
                if (b == 0) pt_shade<true, true><<<blocks, SH_THREADS, SH_SMEM_BYTES, st>>>(p);
                else pt_shade<false, true><<<blocks, SH_THREADS, SH_SMEM_BYTES, st>>>(p);
            }
            else if (b == 0) pt_shade<true><<<blocks, SH_THREADS, SH_SMEM_BYTES, st>>>(p);
            else pt_shade<false><<<blocks, SH_THREADS, SH_SMEM_BYTES, st>>>(p);
        }
        h->launches += 2;
        mark();
        h->cur = nxt;
        if (sort && b + 1 < h->depth) {
            const int nb = std::max(h->nmaterials, 1), srt = (h->cur + 1) % 3;
            sort_hist<<<h->sort_blocks, SORT_TILE, nb * sizeof(int), st>>>(h->d_keys, h->d_counts + b + 1, nb, h->sort_blocks, h->d_hist);
            sort_scan<<<1, 1024, 0, st>>>(h->d_hist, nb * h->sort_blocks);
            sort_scatter<<<h->sort_blocks, SORT_TILE, 0, st>>>(h->d_keys, h->d_counts + b + 1, nb, h->sort_blocks, h->d_hist, h->d_paths[h->cur], h->d_paths[srt]);
            h->launches += 3;
            h->cur = srt;
        }
    }
    if (last >= h->depth) {
        h->final_buf = h->cur;
        if (h->profiling) h->timed_launches = h->nmark - 1;
    }
    CUDA_TRY(cudaGetLastError());
    return PTD_OK;
}

#define PT_NOT_INFLIGHT(h, what) do { if ((h)->fr_waited < (h)->fr_submitted) PTD_FAIL(PTD_ERR_STATE, what ": frame(s) submitted with ptd_frame_submit are still in flight - call ptd_frame_wait first"); } while (0)
extern "C" ptd_status ptd_pt_render(ptd_pt* h, const ptd_camera* cam, int iter, float* gbuf, void* stream_) {
    if (!h || iter < 1) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_render: bad argument");
    PT_NOT_INFLIGHT(h, "ptd_pt_render");
    return pt_run(h, cam, iter, gbuf, (cudaStream_t)stream_, 0, h->depth);
}

// Same-process strip group (tests; single-process multi-GPU): bounce by bounce over the strips in rank order, so that on one
// GPU - where the strips' kernels run one after the other - every live count a kernel waits for has already been mailed.
extern "C" ptd_status ptd_pt_render_group(ptd_pt** hs, int n, const ptd_camera* cam, int iter, float* const* gbufs, void* const* streams) {
    if (!hs || n < 1 || iter < 1) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_render_group: bad argument");
    for (int b = 0; b < hs[0]->depth; ++b)
        for (int i = 0; i < n; ++i) {
            ptd_status rc = pt_run(hs[i], cam, iter, gbufs ? gbufs[i] : nullptr, streams ? (cudaStream_t)streams[i] : nullptr, b, b + 1);
            if (rc != PTD_OK) return rc;
        }
    return PTD_OK;
}

struct ptd_pt_strip_info { unsigned char ipc[64]; unsigned long long mail; int pid_tag, device, row0, rows, W, H, depth, reserved; };
extern "C" int ptd_pt_strip_info_size(void) { return (int)sizeof(ptd_pt_strip_info); }
extern "C" ptd_status ptd_pt_strip_export(ptd_pt* h, void* info_out, int capacity) {
    if (!h || !info_out || capacity < (int)sizeof(ptd_pt_strip_info)) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_strip_export: need a buffer of %zu bytes", sizeof(ptd_pt_strip_info));
    CUDA_TRY(cudaSetDevice(h->device));
    ptd_pt_strip_info info;
    memset(&info, 0, sizeof info);
    cudaIpcMemHandle_t ipc;
    if (cudaIpcGetMemHandle(&ipc, h->d_mail) == cudaSuccess) memcpy(info.ipc, &ipc, sizeof ipc);
    else cudaGetLastError();
    info.mail = (unsigned long long)(uintptr_t)h->d_mail;
    info.pid_tag = (int)getpid(); info.device = h->device; info.row0 = h->row0; info.rows = h->rows; info.W = h->W; info.H = h->H; info.depth = h->depth;
    memcpy(info_out, &info, sizeof info);
    return PTD_OK;
}
// infos: the nranks exported blobs in strip order (top strip first), back to back; my_rank: this handle's position.
extern "C" ptd_status ptd_pt_strip_connect(ptd_pt* h, const void* infos, int nranks, int my_rank) {
    if (!h || !infos || nranks < 1 || nranks > PT_MAX_RANKS || my_rank < 0 || my_rank >= nranks) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_strip_connect: bad argument (at most %d strips)", PT_MAX_RANKS);
    CUDA_TRY(cudaSetDevice(h->device));
    const ptd_pt_strip_info* in = (const ptd_pt_strip_info*)infos;
    int row = 0;
    for (int r = 0; r < nranks; ++r) {
        ptd_pt_strip_info info;
        memcpy(&info, &in[r], sizeof info);
        if (info.W != h->W || info.H != h->H || info.depth != h->depth || info.row0 != row) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_strip_connect: strip %d does not continue the frame at row %d", r, row);
        row += info.rows;
        if (r == my_rank && (info.row0 != h->row0 || info.rows != h->rows)) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_strip_connect: blob %d is not this handle's", r);
    }
    if (row != h->H) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_strip_connect: the strips cover %d of %d rows", row, h->H);
    for (int r = 0; r < PT_MAX_RANKS; ++r) {
        if (h->peer_mail[r] && h->peer_ipc[r]) cudaIpcCloseMemHandle(h->peer_mail[r]);
        h->peer_mail[r] = nullptr; h->peer_ipc[r] = false;
    }
    for (int r = my_rank + 1; r < nranks; ++r) {                      // we only ever write to the strips below us
        ptd_pt_strip_info info;
        memcpy(&info, &in[r], sizeof info);
        if (info.pid_tag == (int)getpid()) {
            if (info.device != h->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(info.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ptd_set_error("ptd_pt_strip_connect: no peer access %d -> %d: %s", h->device, info.device, cudaGetErrorString(e)); return PTD_ERR_CUDA; }
                cudaGetLastError();
            }
            h->peer_mail[r] = (unsigned long long*)(uintptr_t)info.mail;
        } else {
            cudaIpcMemHandle_t ipc;
            memcpy(&ipc, info.ipc, sizeof ipc);
            void* q = nullptr;
            CUDA_TRY(cudaIpcOpenMemHandle(&q, ipc, cudaIpcMemLazyEnablePeerAccess));
            h->peer_mail[r] = (unsigned long long*)q; h->peer_ipc[r] = true;
        }
    }
    h->rank = my_rank; h->nranks = nranks;
    return PTD_OK;
}

extern "C" ptd_status ptd_frame_wait(ptd_pt* h);
extern "C" ptd_status ptd_pt_render_host(ptd_pt* h, const ptd_camera* cam, int iter, float* host_tensor) {
    if (!h || !host_tensor || iter < 1) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_render_host: bad argument");
    while (h->fr_waited < h->fr_submitted) { ptd_status rc_ = ptd_frame_wait(h); if (rc_ != PTD_OK) return rc_; }   // frames submitted asynchronously complete first
    CUDA_TRY(cudaSetDevice(h->device));
    if (!h->host_stream[0]) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->host_stream[0], cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&h->host_stream[1], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->host_event, cudaEventDisableTiming));
    }
    // The normal / depth / albedo planes (7 of the 10) are final after the first bounce (pathtrace.cu:295-304, :379-387): their copy
    // to the caller's host_tensor (pathtrace.cu:525) overlaps bounces 1..depth-1; only the radiance planes wait for the last bounce.
    const size_t plane = sizeof(float) * (size_t)h->Pfull;
    ptd_status rc = pt_run(h, cam, iter, nullptr, h->host_stream[0], 0, 1);
    if (rc != PTD_OK) return rc;
    const bool early = iter == 1;                                       // later iterations do not rewrite those planes at all
    CUDA_TRY(cudaEventRecord(h->host_event, h->host_stream[0]));
    CUDA_TRY(cudaStreamWaitEvent(h->host_stream[1], h->host_event, 0));
    CUDA_TRY(cudaMemcpyAsync(host_tensor + 3 * (size_t)h->Pfull, h->d_gbuf_own + 3 * (size_t)h->Pfull, 7 * plane, cudaMemcpyDeviceToHost, h->host_stream[1]));
    (void)early;
    rc = pt_run(h, cam, iter, nullptr, h->host_stream[0], 1, h->depth);
    if (rc != PTD_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(host_tensor, h->d_gbuf_own, 3 * plane, cudaMemcpyDeviceToHost, h->host_stream[0]));
    CUDA_TRY(cudaStreamSynchronize(h->host_stream[1]));
    CUDA_TRY(cudaStreamSynchronize(h->host_stream[0]));
    return PTD_OK;
}

// One frame of the reference's runCuda() body (main.cpp:143-158: pathtraceInit / pathtrace / network_prediction) as ONE blocking
// call: path trace -> denoise on the device, host copies only where the caller wants them.  Compared with ptd_pt_render_host +
// ptd_dn_forward_host (the two reference call sites taken one by one) the 40*P-byte G-buffer is never uploaded again, and its
// download (host_tensor, optional) overlaps the remaining bounces and the denoiser.
extern "C" ptd_status ptd_frame_host(ptd_pt* h, ptd_dn* dn, const ptd_camera* cam, int iter, int reset_hidden, float* host_tensor, float* rgb_host) {
    if (!h || !dn || !rgb_host || iter < 1) PTD_FAIL(PTD_ERR_ARG, "ptd_frame_host: bad argument");
    while (h->fr_waited < h->fr_submitted) { ptd_status rc_ = ptd_frame_wait(h); if (rc_ != PTD_OK) return rc_; }   // frames submitted asynchronously complete first
    int dn_device = 0, dn_H = 0, dn_W = 0, dn_strip = 0;
    ptd_dn_describe(dn, &dn_device, &dn_H, &dn_W, &dn_strip);
    if (h->nranks > 1 || h->rows != h->H || dn_strip) PTD_FAIL(PTD_ERR_UNSUPPORTED, "ptd_frame_host: row-strip handles take device pointers (ptd_pt_render + ptd_dn_forward)");
    if (dn_device != h->device || dn_H != h->H || dn_W != h->W) PTD_FAIL(PTD_ERR_ARG, "ptd_frame_host: the denoiser handle is for %dx%d on device %d, the path tracer for %dx%d on device %d", dn_W, dn_H, dn_device, h->W, h->H, h->device);
    CUDA_TRY(cudaSetDevice(h->device));
    if (!h->host_stream[0]) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->host_stream[0], cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&h->host_stream[1], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->host_event, cudaEventDisableTiming));
    }
    const size_t plane = sizeof(float) * (size_t)h->Pfull;
    if (!h->d_frame_rgb) CUDA_TRY(cudaMalloc((void**)&h->d_frame_rgb, 3 * plane));
    ptd_status rc = pt_run(h, cam, iter, nullptr, h->host_stream[0], 0, 1);
    if (rc != PTD_OK) return rc;
    if (host_tensor) {                                                  // normal / depth / albedo planes are final after bounce 0
        CUDA_TRY(cudaEventRecord(h->host_event, h->host_stream[0]));
        CUDA_TRY(cudaStreamWaitEvent(h->host_stream[1], h->host_event, 0));
        CUDA_TRY(cudaMemcpyAsync(host_tensor + 3 * (size_t)h->Pfull, h->d_gbuf_own + 3 * (size_t)h->Pfull, 7 * plane, cudaMemcpyDeviceToHost, h->host_stream[1]));
    }
    rc = pt_run(h, cam, iter, nullptr, h->host_stream[0], 1, h->depth);
    if (rc != PTD_OK) return rc;
    if (host_tensor) {                                                  // the radiance planes travel while the denoiser runs
        CUDA_TRY(cudaEventRecord(h->host_event, h->host_stream[0]));
        CUDA_TRY(cudaStreamWaitEvent(h->host_stream[1], h->host_event, 0));
        CUDA_TRY(cudaMemcpyAsync(host_tensor, h->d_gbuf_own, 3 * plane, cudaMemcpyDeviceToHost, h->host_stream[1]));
    }
    rc = ptd_dn_forward(dn, h->d_gbuf_own, h->d_frame_rgb, reset_hidden, h->host_stream[0]);
    if (rc != PTD_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(rgb_host, h->d_frame_rgb, 3 * plane, cudaMemcpyDeviceToHost, h->host_stream[0]));
    CUDA_TRY(cudaStreamSynchronize(h->host_stream[1]));
    CUDA_TRY(cudaStreamSynchronize(h->host_stream[0]));
    return PTD_OK;
}

// ---- the same frame, asynchronously: submit frame k + 1, then wait for frame k ----------------------------------------------------
// ptd_frame_submit enqueues a whole frame - path trace on one stream, denoiser + the copy of the denoised frame on a second, the
// optional G-buffer copy on a third - into one of two buffer slots and returns at once; ptd_frame_wait blocks until the OLDEST
// submitted frame has reached the caller's host buffers.  With one frame in flight while the next is submitted, the path trace of
// frame k + 1 overlaps the denoiser and the PCIe copies of frame k (the reference's loop is strictly serial, main.cpp:120-168).
// PTD_FRAME_SM_SPLIT=<n>: the denoiser stream gets its own partition of >= n SMs, the path-trace stream the rest (CUDA green contexts).
// Why: the two halves of consecutive frames are meant to overlap, but a conv CTA needs ~61 K of an SM's 64 K registers and a trace block 16 K,
// so an SM can hold one kind or the other, never both.  When both streams have work the block scheduler spreads each grid over all SMs and the
// two kernels end up waiting for each other's blocks to drain; with a strip's small grids (multi-GPU mode) neither kernel fills the GPU and the
// partition lets both run at once.  Untiled frames fill the GPU either way (the split is work-conserving there: no gain).
#define FRAME_SPLIT_MIN_RANKS 2
#define FRAME_SPLIT_DN_SMS 32
#define CU_TRY_DRV(call) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) PTD_FAIL(PTD_ERR_CUDA, "%s failed with CUresult %d", #call, (int)r_); } while (0)
template <typename F> static F drv_entry(const char* name) {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return (F)fp;
}
static void frame_green_destroy(void* green_ctx) {
    typedef CUresult (*F_destroy)(CUgreenCtx);
    if (F_destroy destroy = drv_entry<F_destroy>("cuGreenCtxDestroy")) destroy((CUgreenCtx)green_ctx);
}
static ptd_status frame_sm_split(ptd_pt* h, int sm_dn) {
    typedef CUresult (*F_getres)(CUdevice, CUdevResource*, CUdevResourceType);
    typedef CUresult (*F_split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    typedef CUresult (*F_desc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    typedef CUresult (*F_create)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*F_stream)(CUstream*, CUgreenCtx, unsigned int, int);
    typedef CUresult (*F_devget)(CUdevice*, int);
    F_getres getres = drv_entry<F_getres>("cuDeviceGetDevResource");
    F_split splitf = drv_entry<F_split>("cuDevSmResourceSplitByCount");
    F_desc desc = drv_entry<F_desc>("cuDevResourceGenerateDesc");
    F_create create = drv_entry<F_create>("cuGreenCtxCreate");
    F_stream mkstream = drv_entry<F_stream>("cuGreenCtxStreamCreate");
    F_devget devget = drv_entry<F_devget>("cuDeviceGet");
    if (!getres || !splitf || !desc || !create || !mkstream || !devget) PTD_FAIL(PTD_ERR_UNSUPPORTED, "PTD_FRAME_SM_SPLIT: this driver has no green contexts");
    CUdevice dev; CU_TRY_DRV(devget(&dev, h->device));
    CUdevResource all, part, rest;
    CU_TRY_DRV(getres(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    unsigned int groups = 1;
    CU_TRY_DRV(splitf(&part, &groups, &all, &rest, 0, (unsigned)sm_dn));
    if (groups != 1 || rest.sm.smCount < 8) PTD_FAIL(PTD_ERR_ARG, "PTD_FRAME_SM_SPLIT=%d: cannot split %u SMs that way", sm_dn, all.sm.smCount);
    CUdevResourceDesc d_dn, d_pt;
    CU_TRY_DRV(desc(&d_dn, &part, 1));
    CU_TRY_DRV(desc(&d_pt, &rest, 1));
    CUgreenCtx g_dn, g_pt;
    CU_TRY_DRV(create(&g_pt, d_pt, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CU_TRY_DRV(create(&g_dn, d_dn, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream s_pt, s_dn;
    CU_TRY_DRV(mkstream(&s_pt, g_pt, CU_STREAM_NON_BLOCKING, 0));
    CU_TRY_DRV(mkstream(&s_dn, g_dn, CU_STREAM_NON_BLOCKING, 0));
    h->fr_stream[0] = (cudaStream_t)s_pt; h->fr_stream[1] = (cudaStream_t)s_dn;
    h->fr_green[0] = (void*)g_pt; h->fr_green[1] = (void*)g_dn;
    h->fr_sm_pt = (int)rest.sm.smCount; h->fr_sm_dn = (int)part.sm.smCount;
    return PTD_OK;
}
// At most FRAME_SLOTS frames may be in flight; frames complete in submission order; the recurrent state is carried in that order.
static ptd_status frame_ring_init(ptd_pt* h) {
    if (h->fr_stream[0]) return PTD_OK;
    const size_t plane = sizeof(float) * (size_t)h->Pfull;
    // SM partition of the two streams (see frame_sm_split): PTD_FRAME_SM_SPLIT=<SMs of the denoiser> (0 = none); by default 32 for the strips
    // of a frame tiled over several GPUs, where neither half fills the GPU (C3, 2xf16: N = 2 340 -> 359 frames/s, N = 4 514 -> 576, N = 8 725 -> 821;
    // with 16 / 40 / 64 SMs: 600 at N = 8, 549 / 471 at N = 4), none for untiled frames (work-conserving there: 32 SMs 209.5 vs 210.8 frames/s, 48 SMs 185.6).
    const bool two_streams = (h->nranks == 1 && h->rows == h->H) || (h->flags & PTD_PT_GATED_MAIL);
    const char* e = getenv("PTD_FRAME_SM_SPLIT");
    const int split = e ? atoi(e) : (h->nranks >= FRAME_SPLIT_MIN_RANKS ? FRAME_SPLIT_DN_SMS : 0);
    if (split > 0 && two_streams) {
        ptd_status rc = frame_sm_split(h, split);
        if (rc != PTD_OK && e) return rc;                               // asked for explicitly: fail loudly; the default falls back to the shared GPU
    }
    for (int i = 0; i < 3; ++i) if (!h->fr_stream[i]) CUDA_TRY(cudaStreamCreateWithFlags(&h->fr_stream[i], cudaStreamNonBlocking));
    for (int i = 0; i < FRAME_SLOTS; ++i) {
        CUDA_TRY(cudaMalloc((void**)&h->fr_gbuf[i], 10 * plane));
        CUDA_TRY(cudaMalloc((void**)&h->fr_rgb[i], 3 * plane));
        CUDA_TRY(cudaMemset(h->fr_gbuf[i], 0, 10 * plane));
        CUDA_TRY(cudaEventCreateWithFlags(&h->fr_ev_pt[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->fr_ev_done[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->fr_ev_gcopy[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->fr_ev_rcopy[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaDeviceSynchronize());                                  // the memsets ran on the legacy stream
    return PTD_OK;
}
extern "C" ptd_status ptd_frame_submit(ptd_pt* h, ptd_dn* dn, const ptd_camera* cam, int iter, int reset_hidden, float* host_tensor, float* rgb_host) {
    if (!h || !dn || iter < 1) PTD_FAIL(PTD_ERR_ARG, "ptd_frame_submit: bad argument");
    // the first-hit planes 3..9 are written by iteration 1 only (pathtrace.cu:295, :379) and live in the slot's G-buffer: accumulating further
    // iterations would need them carried from slot to slot - use ptd_frame_host / ptd_pt_render for iter > 1
    if (iter != 1) PTD_FAIL(PTD_ERR_UNSUPPORTED, "ptd_frame_submit: only iter == 1 (one sample per pixel per frame, as runCuda() renders); use ptd_frame_host for accumulation");
    int dn_device = 0, dn_H = 0, dn_W = 0, dn_strip = 0;
    ptd_dn_describe(dn, &dn_device, &dn_H, &dn_W, &dn_strip);
    const bool strip = h->nranks > 1 || h->rows != h->H;
    if (strip != (dn_strip != 0)) PTD_FAIL(PTD_ERR_ARG, "ptd_frame_submit: path tracer and denoiser handles must both cover the frame or both be row strips");
    if (dn_device != h->device || dn_H != h->H || dn_W != h->W) PTD_FAIL(PTD_ERR_ARG, "ptd_frame_submit: the denoiser handle is for %dx%d on device %d, the path tracer for %dx%d on device %d", dn_W, dn_H, dn_device, h->W, h->H, h->device);
    if (h->fr_submitted - h->fr_waited >= FRAME_SLOTS) PTD_FAIL(PTD_ERR_STATE, "ptd_frame_submit: %d frames are in flight - call ptd_frame_wait first", FRAME_SLOTS);
    CUDA_TRY(cudaSetDevice(h->device));
    ptd_status rc = frame_ring_init(h);
    if (rc != PTD_OK) return rc;
    const int i = (int)(h->fr_submitted % FRAME_SLOTS);                 // this slot's previous frame (FRAME_SLOTS submissions ago) has been waited for
    const size_t plane = sizeof(float) * (size_t)h->Pfull;
    // Row strips: a frame's kernels wait for other GPUs (halo rows, live counts).  With the path trace of frame k + 1 on its own stream,
    // a kernel that spins while it holds an SM can starve the kernel it waits for; PTD_PT_GATED_MAIL moves the path tracer's only wait
    // into a one-warp gate kernel, so only handles created with it overlap the two - the others run a frame's
    // path trace and denoiser on one stream (submission is still asynchronous, the host copies still overlap the next frame).
    const bool two_streams = !strip || (h->flags & PTD_PT_GATED_MAIL);
    cudaStream_t s_pt = h->fr_stream[0], s_dn = two_streams ? h->fr_stream[1] : h->fr_stream[0], s_cp = h->fr_stream[2];
    if (h->fr_timer_armed) { CUDA_TRY(cudaEventRecord(h->fr_ev_t0, s_pt)); h->fr_timer_armed = false; }
    if (two_streams && h->fr_submitted >= FRAME_SLOTS) CUDA_TRY(cudaStreamWaitEvent(s_pt, h->fr_ev_done[i], 0));   // the slot's G-buffer is free once its previous frame was denoised
#ifdef PTD_FRAME_SPANS
    for (int e = 0; e < 4; ++e) if (!h->sp_ev[i][e]) cudaEventCreate(&h->sp_ev[i][e]);
    cudaEventRecord(h->sp_ev[i][0], s_pt);
#endif
    rc = pt_run(h, cam, iter, h->fr_gbuf[i], s_pt, 0, h->depth);
#ifdef PTD_FRAME_SPANS
    cudaEventRecord(h->sp_ev[i][1], s_pt);
#endif
    if (rc != PTD_OK) { cudaStreamSynchronize(s_pt); return rc; }
    CUDA_TRY(cudaEventRecord(h->fr_ev_pt[i], s_pt));
    // the rows this handle renders: the whole frame, or its strip of every plane ([planes][H][W] on both sides)
    const size_t row_off = (size_t)h->row0 * h->W, row_bytes = sizeof(float) * (size_t)h->rows * h->W;
    h->fr_has_gcopy[i] = host_tensor != nullptr;
    if (host_tensor) {
        CUDA_TRY(cudaStreamWaitEvent(s_cp, h->fr_ev_pt[i], 0));
        CUDA_TRY(cudaMemcpy2DAsync(host_tensor + row_off, plane, h->fr_gbuf[i] + row_off, plane, row_bytes, 10, cudaMemcpyDeviceToHost, s_cp));
        CUDA_TRY(cudaEventRecord(h->fr_ev_gcopy[i], s_cp));
    }
    if (two_streams) CUDA_TRY(cudaStreamWaitEvent(s_dn, h->fr_ev_pt[i], 0));
#ifdef PTD_FRAME_SPANS
    cudaEventRecord(h->sp_ev[i][2], s_dn);
#endif
    ptd_dn_set_sm_limit(dn, two_streams ? h->fr_sm_dn : 0);
    rc = ptd_dn_forward_frame(dn, h->fr_gbuf[i], h->fr_rgb[i], reset_hidden, (void*)s_dn);
#ifdef PTD_FRAME_SPANS
    cudaEventRecord(h->sp_ev[i][3], s_dn);
#endif
    if (rc != PTD_OK) { cudaStreamSynchronize(s_pt); cudaStreamSynchronize(s_dn); return rc; }
    CUDA_TRY(cudaEventRecord(h->fr_ev_done[i], s_dn));
    h->fr_has_rcopy[i] = rgb_host != nullptr;
    if (rgb_host) {                                                     // on the copy stream: the denoiser stream goes straight on to the next frame
        CUDA_TRY(cudaStreamWaitEvent(s_cp, h->fr_ev_done[i], 0));
        CUDA_TRY(cudaMemcpy2DAsync(rgb_host + row_off, plane, h->fr_rgb[i] + row_off, plane, row_bytes, 3, cudaMemcpyDeviceToHost, s_cp));
        CUDA_TRY(cudaEventRecord(h->fr_ev_rcopy[i], s_cp));
    }
    h->fr_last_dn = s_dn;
    h->fr_dn[i] = dn;
    ptd_dn_mark_inflight(dn, +1);
    h->fr_submitted += 1;
    return PTD_OK;
}
extern "C" int ptd_frame_slots(void) { return FRAME_SLOTS; }
extern "C" ptd_status ptd_frame_wait(ptd_pt* h) {
    if (!h) PTD_FAIL(PTD_ERR_ARG, "ptd_frame_wait: null handle");
    if (h->fr_waited == h->fr_submitted) PTD_FAIL(PTD_ERR_STATE, "ptd_frame_wait: no frame in flight");
    CUDA_TRY(cudaSetDevice(h->device));
    const int i = (int)(h->fr_waited % FRAME_SLOTS);
    h->fr_waited += 1;                                                  // whatever happens below, this frame is no longer in flight
    if (h->fr_dn[i]) { ptd_dn_mark_inflight(h->fr_dn[i], -1); h->fr_dn[i] = nullptr; }
    CUDA_TRY(cudaEventSynchronize(h->fr_ev_done[i]));
    if (h->fr_has_gcopy[i]) CUDA_TRY(cudaEventSynchronize(h->fr_ev_gcopy[i]));
    if (h->fr_has_rcopy[i]) CUDA_TRY(cudaEventSynchronize(h->fr_ev_rcopy[i]));
    return PTD_OK;
}
// Device time of a run of submitted frames: op 0 arms the timer (the next ptd_frame_submit records the start event on the path-trace
// stream before its first launch); op 1 records the stop event behind the last submitted frame's denoiser, waits for it and returns
// the milliseconds in between (every frame in flight is complete afterwards, but still has to be taken with ptd_frame_wait).
extern "C" ptd_status ptd_frame_timer(ptd_pt* h, int op, float* ms) {
    if (!h || (op == 1 && !ms)) PTD_FAIL(PTD_ERR_ARG, "ptd_frame_timer: bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    ptd_status rc = frame_ring_init(h);
    if (rc != PTD_OK) return rc;
    if (!h->fr_ev_t0) { CUDA_TRY(cudaEventCreate(&h->fr_ev_t0)); CUDA_TRY(cudaEventCreate(&h->fr_ev_t1)); }
    if (op == 0) { h->fr_timer_armed = true; return PTD_OK; }
    if (h->fr_timer_armed || !h->fr_last_dn) PTD_FAIL(PTD_ERR_STATE, "ptd_frame_timer: no frame was submitted since the timer was armed");
    CUDA_TRY(cudaEventRecord(h->fr_ev_t1, h->fr_last_dn));
    CUDA_TRY(cudaEventSynchronize(h->fr_ev_t1));
    CUDA_TRY(cudaEventElapsedTime(ms, h->fr_ev_t0, h->fr_ev_t1));
    return PTD_OK;
}

#ifdef PTD_FRAME_SPANS
// debug build only: [PT start, PT end, DN start, DN end] of the last two submitted frames (older first), ms since the older frame's PT start
extern "C" int ptd_debug_frame_spans(ptd_pt* h, float* out8) {
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    const int older = (int)((h->fr_submitted + FRAME_SLOTS - 2) % FRAME_SLOTS);
    for (int f = 0; f < 2; ++f)
        for (int e = 0; e < 4; ++e) cudaEventElapsedTime(&out8[f * 4 + e], h->sp_ev[older][0], h->sp_ev[(older + f) % FRAME_SLOTS][e]);
    return 0;
}
#endif
extern "C" ptd_status ptd_pt_export_rgba8(ptd_pt* h, int iter, unsigned char* pbo, void* stream_) {
    if (!h || !pbo || iter < 1) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_export_rgba8: bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    dim3 b(8, 8), g((h->W + 7) / 8, (h->rows + 7) / 8);
    export_rgba8<<<g, b, 0, (cudaStream_t)stream_>>>(h->d_image, h->W, h->rows, iter, (uchar4*)pbo + (size_t)h->row0 * h->W);
    CUDA_TRY(cudaGetLastError());
    return PTD_OK;
}

extern "C" ptd_status ptd_pt_live_counts(ptd_pt* h, int* counts, int capacity, int* bounces_run) {
    if (!h || !counts || capacity < h->depth) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_live_counts: need capacity >= trace depth %d", h ? h->depth : 0);
    CUDA_TRY(cudaSetDevice(h->device));
    PT_NOT_INFLIGHT(h, "ptd_pt_live_counts");
    CUDA_TRY(cudaDeviceSynchronize());                                  // the last render may have run on a non-blocking stream
    std::vector<int> c(h->depth + 1);
    CUDA_TRY(cudaMemcpy(c.data(), h->d_counts, sizeof(int) * (h->depth + 1), cudaMemcpyDeviceToHost));
    c[0] = h->P;
    int run = 0;
    for (int b = 0; b < h->depth; ++b) { counts[b] = c[b]; if (c[b] > 0) run = b + 1; }
    if (bounces_run) *bounces_run = run;
    return PTD_OK;
}
static ptd_status bounce_count(ptd_pt* h, int bounce, int* n) {
    if (bounce < 0 || bounce >= h->depth) PTD_FAIL(PTD_ERR_ARG, "bounce %d out of range", bounce);
    if (bounce == 0) { *n = h->P; return PTD_OK; }
    CUDA_TRY(cudaMemcpy(n, h->d_counts + bounce, sizeof(int), cudaMemcpyDeviceToHost));
    return PTD_OK;
}
extern "C" ptd_status ptd_pt_dump_paths(ptd_pt* h, int bounce, ptd_path_segment* host, int capacity, int* n) {
    if (!h || !host || !n) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_dump_paths: null argument");
    if (!h->d_trace_paths) PTD_FAIL(PTD_ERR_STATE, "ptd_pt_dump_paths: handle was created without PTD_PT_TRACE");
    CUDA_TRY(cudaSetDevice(h->device));
    PT_NOT_INFLIGHT(h, "ptd_pt_dump_paths");
    CUDA_TRY(cudaDeviceSynchronize());                                  // the last render may have run on a non-blocking stream
    ptd_status rc = bounce_count(h, bounce, n);
    if (rc != PTD_OK) return rc;
    if (*n > capacity) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_dump_paths: capacity %d < %d", capacity, *n);
    CUDA_TRY(cudaMemcpy(host, h->d_trace_paths + (size_t)bounce * h->P, sizeof(ptd_path_segment) * (size_t)*n, cudaMemcpyDeviceToHost));
    return PTD_OK;
}
extern "C" ptd_status ptd_pt_dump_intersections(ptd_pt* h, int bounce, ptd_intersection* host, int capacity, int* n) {
    if (!h || !host || !n) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_dump_intersections: null argument");
    if (!h->d_trace_isx) PTD_FAIL(PTD_ERR_STATE, "ptd_pt_dump_intersections: handle was created without PTD_PT_TRACE");
    CUDA_TRY(cudaSetDevice(h->device));
    PT_NOT_INFLIGHT(h, "ptd_pt_dump_intersections");
    CUDA_TRY(cudaDeviceSynchronize());                                  // the last render may have run on a non-blocking stream
    ptd_status rc = bounce_count(h, bounce, n);
    if (rc != PTD_OK) return rc;
    if (*n > capacity) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_dump_intersections: capacity %d < %d", capacity, *n);
    CUDA_TRY(cudaMemcpy(host, h->d_trace_isx + (size_t)bounce * h->P, sizeof(ptd_intersection) * (size_t)*n, cudaMemcpyDeviceToHost));
    return PTD_OK;
}
extern "C" ptd_status ptd_pt_dump_final_paths(ptd_pt* h, ptd_path_segment* host, int capacity) {
    if (!h || !host || capacity < h->P) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_dump_final_paths: bad argument");
    if (!h->d_dead) PTD_FAIL(PTD_ERR_STATE, "ptd_pt_dump_final_paths: handle was created without PTD_PT_KEEP_TERMINATED");
    CUDA_TRY(cudaSetDevice(h->device));
    PT_NOT_INFLIGHT(h, "ptd_pt_dump_final_paths");
    CUDA_TRY(cudaDeviceSynchronize());                                  // the last render may have run on a non-blocking stream
    CUDA_TRY(cudaMemcpy(host, h->d_dead, sizeof(ptd_path_segment) * (size_t)h->P, cudaMemcpyDeviceToHost));
    return PTD_OK;
}
extern "C" ptd_status ptd_pt_dump_image(ptd_pt* h, float* host_rgb) {
    if (!h || !host_rgb) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_dump_image: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    PT_NOT_INFLIGHT(h, "ptd_pt_dump_image");
    CUDA_TRY(cudaDeviceSynchronize());                                  // the last render may have run on a non-blocking stream
    CUDA_TRY(cudaMemcpy(host_rgb, h->d_image, sizeof(float) * 3 * (size_t)h->P, cudaMemcpyDeviceToHost));
    return PTD_OK;
}
extern "C" ptd_status ptd_pt_bvh_stats(const ptd_pt* h, int* nodes, int* leaves, int* max_leaf, int* max_depth) {
    if (!h) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_bvh_stats: null handle");
    if (nodes) *nodes = h->bvh_nodes;
    if (leaves) *leaves = h->bvh_leaves;
    if (max_leaf) *max_leaf = h->bvh_max_leaf;
    if (max_depth) *max_depth = h->bvh_max_depth;
    return PTD_OK;
}
extern "C" int ptd_pt_launches_last_render(const ptd_pt* h) { return h ? h->launches : 0; }
extern "C" ptd_status ptd_pt_profile(ptd_pt* h, int enable) {
    if (!h) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_profile: null handle");
    h->profiling = enable != 0;
    h->timed_launches = 0;
    return PTD_OK;
}
extern "C" ptd_status ptd_pt_launch_times(ptd_pt* h, float* ms, int capacity, int* n) {
    if (!h || !ms || !n) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_launch_times: null argument");
    if (h->timed_launches <= 0) PTD_FAIL(PTD_ERR_STATE, "ptd_pt_launch_times: no profiled render has run");
    if (capacity < h->timed_launches) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_launch_times: capacity %d < %d", capacity, h->timed_launches);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventSynchronize(h->events[h->timed_launches]));
    for (int i = 0; i < h->timed_launches; ++i) CUDA_TRY(cudaEventElapsedTime(&ms[i], h->events[i], h->events[i + 1]));
    *n = h->timed_launches;
    return PTD_OK;
}
