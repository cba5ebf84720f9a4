// HP-1: the 1-spp path-trace iteration, B200-native.
//
// Replaces pathtraceInit / pathtrace / pathtraceFree (Inference/src/pathtrace.cu:96-145, :422-528).
// Where the reference runs, per bounce, memset + computeIntersections + cudaDeviceSynchronize + shadeMaterial +
// thrust::partition (temp malloc, 3 kernels, D2H of the count) and after the loop finalGather + copy_data + a
// 40*P-byte D2H, this file runs ONE fused kernel per bounce and nothing after the loop:
//
//   pt_bounce<FIRST>:  [ray generation at bounce 0 | coalesced smem-staged load of the 44-byte PathSegment tile]
//                      -> nearest hit over the geoms (shared-memory copy) and the mesh (BVH, tie-break exact)
//                      -> shade / scatter (same RNG stream: seed = hash(iter, compacted index, remainingBounces))
//                      -> G-buffer planes written in place (normal/depth/albedo at bounce 0, radiance at termination)
//                      -> stable stream compaction: warp-ballot ranks + block scan + decoupled look-back across tiles,
//                         survivors staged in shared memory and stored coalesced; the live count stays on the device.
//
// The PathSegment array keeps the reference's 44-byte AoS layout in HBM (so parity dumps are plain copies and the
// algorithmic bytes are the survey's 44 B read + 44 B written per live path per bounce); coalescing comes from the
// shared-memory staging, not from a layout change.  No host synchronisation happens inside a frame.
#include <cuda_runtime.h>
#include <cstring>
#include <vector>
#include "ptd_internal.h"
#include "pt_math.cuh"

#define PT_BLOCK 128
#define PT_WORDS 11                 // sizeof(PathSegment) / 4
#define PT_STACK 64

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ptd_set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return PTD_ERR_CUDA; } } while (0)

struct PtKernelParams {
    const ptd_geom* geoms; const ptd_aabb* geom_bounds; int ngeoms; int geoms_in_smem;
    const ptd_material* materials; int nmaterials;
    const ptd_face* faces; int nfaces;
    const float4* nodes; const float4* tris; int use_bvh;
    ptd_aabb mesh_box;
    ptd_camera cam;
    int W, P, iter, bounce, trace_depth;
    const ptd_path_segment* src; ptd_path_segment* dst; ptd_path_segment* dead;
    int* counts;                    // counts[b] = live paths entering bounce b
    unsigned long long* status;     // decoupled look-back tile states of this bounce
    int* ticket;                    // dynamic tile id of this bounce
    float* gbuf; float* image;
    int* sort_keys;
    ptd_path_segment* trace_paths; ptd_intersection* trace_isx;
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

// "while-while" traversal with a postponed leaf (Aila & Laine): all lanes of a warp walk interior nodes until every lane holds
// a leaf, then all lanes run triangle tests, so the two divergent code paths stay converged.  One 64-byte fetch gives both
// children's (padded) boxes.  The triangle test itself is the reference's exact expression tree; the BVH only decides WHICH
// faces are tested, and subtrees are skipped only when their entry distance is strictly beyond the best t so far.
#define PT_SENTINEL 0x76543210
__device__ __forceinline__ void traverse_bvh(const float4* __restrict__ nodes, const float4* __restrict__ tris, const Ray ray,
                                             float& t_min, int& best_face, float& bbx, float& bby, bool& hit_face) {
    const float ooeps = 1e-30f;
    const float idx = 1.0f / (fabsf(ray.d.x) > ooeps ? ray.d.x : copysignf(ooeps, ray.d.x));
    const float idy = 1.0f / (fabsf(ray.d.y) > ooeps ? ray.d.y : copysignf(ooeps, ray.d.y));
    const float idz = 1.0f / (fabsf(ray.d.z) > ooeps ? ray.d.z : copysignf(ooeps, ray.d.z));
    const float oodx = ray.o.x * idx, oody = ray.o.y * idy, oodz = ray.o.z * idz;
    int stack[PT_STACK];
    stack[0] = PT_SENTINEL;
    int sp = 0;
    int node = 0, leaf = 0;
    while (node != PT_SENTINEL) {
        bool searching = true;
        while (node >= 0 && node != PT_SENTINEL) {
            const float4 n0 = __ldg(&nodes[4 * node]), n1 = __ldg(&nodes[4 * node + 1]), n2 = __ldg(&nodes[4 * node + 2]);
            const float4 cn = __ldg(&nodes[4 * node + 3]);
            // slabs; the 1e-5 relative slack keeps the (already padded) boxes conservative against the rounding of these products
            const float c0lox = n0.x * idx - oodx, c0hix = n0.y * idx - oodx, c0loy = n0.z * idy - oody, c0hiy = n0.w * idy - oody;
            const float c0loz = n2.x * idz - oodz, c0hiz = n2.y * idz - oodz;
            const float c1lox = n1.x * idx - oodx, c1hix = n1.y * idx - oodx, c1loy = n1.z * idy - oody, c1hiy = n1.w * idy - oody;
            const float c1loz = n2.z * idz - oodz, c1hiz = n2.w * idz - oodz;
            const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.0f));
            const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fmaxf(c0loz, c0hiz));
            const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.0f));
            const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fmaxf(c1loz, c1hiz));
            const float tlim = t_min * 1.00001f;
            const bool h0 = c0min * 0.99999f <= c0max * 1.00001f && c0min * 0.99999f <= tlim;
            const bool h1 = c1min * 0.99999f <= c1max * 1.00001f && c1min * 0.99999f <= tlim;
            const int child0 = __float_as_int(cn.x), child1 = __float_as_int(cn.y);
            if (!h0 && !h1) {
                node = stack[sp--];
            } else {
                node = h0 ? child0 : child1;
                if (h0 && h1) {
                    int other = child1;
                    if (c1min < c0min) { other = node; node = child1; }
                    stack[++sp] = other;
                }
            }
            if (node < 0 && leaf >= 0) {            // first leaf found: postpone it and keep walking
                searching = false;
                leaf = node;
                node = stack[sp--];
            }
            if (!__any_sync(__activemask(), searching)) break;
        }
        while (leaf < 0) {
            const int code = ~leaf, first = code >> 4, count = (code & 15) + 1;
            for (int i = 0; i < count; ++i) {
                const float4 a = __ldg(&tris[3 * (first + i)]), b = __ldg(&tris[3 * (first + i) + 1]), c = __ldg(&tris[3 * (first + i) + 2]);
                float bx, by;
                const float t = triangleParam(V(a.x, a.y, a.z), V(b.x, b.y, b.z), V(c.x, c.y, c.z), ray, bx, by);
                consider_face(t, __float_as_int(a.w), bx, by, t_min, best_face, bbx, bby, hit_face);
            }
            leaf = node;
            if (node < 0) node = stack[sp--];
        }
    }
}

// conservative slab test against a geom's padded world box: false only when the ray certainly misses the geom
__device__ __forceinline__ bool ray_may_hit_box(const Ray& ray, const ptd_aabb& b) {
    const float ooeps = 1e-30f;
    const float idx = 1.0f / (fabsf(ray.d.x) > ooeps ? ray.d.x : copysignf(ooeps, ray.d.x));
    const float idy = 1.0f / (fabsf(ray.d.y) > ooeps ? ray.d.y : copysignf(ooeps, ray.d.y));
    const float idz = 1.0f / (fabsf(ray.d.z) > ooeps ? ray.d.z : copysignf(ooeps, ray.d.z));
    const float x0 = (b.lb.x - ray.o.x) * idx, x1 = (b.ub.x - ray.o.x) * idx;
    const float y0 = (b.lb.y - ray.o.y) * idy, y1 = (b.ub.y - ray.o.y) * idy;
    const float z0 = (b.lb.z - ray.o.z) * idz, z1 = (b.ub.z - ray.o.z) * idz;
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    return !(tn * 0.9999f > tf * 1.0001f);          // NaN compares false -> "may hit"
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

template <bool FIRST>
__global__ void __launch_bounds__(PT_BLOCK) pt_bounce(const PtKernelParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* s_words = reinterpret_cast<uint32_t*>(smem_raw);                     // PT_BLOCK * 11 words staging
    ptd_geom* s_geoms = reinterpret_cast<ptd_geom*>(smem_raw + PT_BLOCK * PT_WORDS * 4);
    __shared__ int s_tile, s_warp_kept[PT_BLOCK / 32], s_excl;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = FIRST ? p.P : p.counts[p.bounce];
    if (tid == 0) s_tile = atomicAdd(p.ticket, 1);
    if (p.geoms_in_smem) {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(p.geoms);
        uint32_t* d = reinterpret_cast<uint32_t*>(s_geoms);
        for (int i = tid; i < p.ngeoms * 62; i += PT_BLOCK) d[i] = __ldg(&g[i]);
        const uint32_t* gb = reinterpret_cast<const uint32_t*>(p.geom_bounds);
        uint32_t* db = reinterpret_cast<uint32_t*>(s_geoms + p.ngeoms);
        for (int i = tid; i < p.ngeoms * 6; i += PT_BLOCK) db[i] = __ldg(&gb[i]);
    }
    __syncthreads();
    const int tile = s_tile;
    const int base = tile * PT_BLOCK;
    if (base >= n) return;
    const int valid = min(PT_BLOCK, n - base);
    const int idx = base + tid;
    const bool active = tid < valid;
    const ptd_geom* geoms = p.geoms_in_smem ? s_geoms : p.geoms;
    const ptd_aabb* gbounds = p.geoms_in_smem ? reinterpret_cast<const ptd_aabb*>(s_geoms + p.ngeoms) : p.geom_bounds;

    // ---- 1. this tile's path segments ---------------------------------------------------------------------
    Ray ray; v3 color; int pixelIndex = 0, rb = 0;
    ray.o = ray.d = color = V(0, 0, 0);
    if (FIRST) {
        if (active) {                                                      // generateRayFromCamera, pathtrace.cu:155-182
            const int x = idx % p.W, y = idx / p.W;
            Rng rng = make_rng(p.iter, idx, 0);                            // stale remainingBounces := 0 (decision D4)
            ray.o = p.cam.position;
            color = V(1.0f, 1.0f, 1.0f);
            float jx = rng_uniform(rng, -0.5f, 0.5f);
            float jy = rng_uniform(rng, -0.5f, 0.5f);
            v3 a = muls(muls(p.cam.right, p.cam.pixelLength_x), fadd(ffma((float)p.cam.res_x, -0.5f, (float)x), jx));
            v3 b = muls(muls(p.cam.up, p.cam.pixelLength_y), fadd(ffma((float)p.cam.res_y, -0.5f, (float)y), jy));
            ray.d = normalize(sub(sub(p.cam.view, a), b));
            pixelIndex = idx;
            rb = p.trace_depth;
        }
    } else {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.src) + (size_t)base * PT_WORDS;
        const int nwords = valid * PT_WORDS;
        if (valid == PT_BLOCK) {                                           // 5632 B tile, 16-B aligned: 352 float4
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(s_words);
            for (int i = tid; i < PT_BLOCK * PT_WORDS / 4; i += PT_BLOCK) d4[i] = s4[i];
        } else {
            for (int i = tid; i < nwords; i += PT_BLOCK) s_words[i] = src[i];
        }
        __syncthreads();
        if (active) {
            const float* w = reinterpret_cast<const float*>(s_words) + tid * PT_WORDS;   // stride 11 words: conflict free
            ray.o = V(w[0], w[1], w[2]); ray.d = V(w[3], w[4], w[5]); color = V(w[6], w[7], w[8]);
            pixelIndex = __float_as_int(w[9]); rb = __float_as_int(w[10]);
        }
        __syncthreads();                                                   // staging buffer is reused for the stores below
    }
    if (p.trace_paths && active) {
        ptd_path_segment ps;
        ps.ray.origin = ray.o; ps.ray.direction = ray.d; ps.color = color; ps.pixelIndex = pixelIndex; ps.remainingBounces = rb;
        p.trace_paths[(size_t)p.bounce * p.P + idx] = ps;
    }

    bool keep = false;
    if (active) {
        // ---- 2. nearest intersection (computeIntersections, pathtrace.cu:200-306) ---------------------------
        float t_min = FLT_MAX;
        v3 ip = V(0, 0, 0), normal = V(0, 0, 0), tip = V(0, 0, 0), tn = V(0, 0, 0);
        int materialid = -1;
        bool outside = true;
        for (int g = 0; g < p.ngeoms; ++g) {
            const ptd_geom* ge = &geoms[g];
            if (!ray_may_hit_box(ray, gbounds[g])) continue;                  // a certain miss leaves t, outside untouched in the reference too
            float t = 0.f;
            if (ge->type == PTD_CUBE) t = boxIntersectionTest(ge, ray, tip, tn, outside);
            else if (ge->type == PTD_SPHERE) t = sphereIntersectionTest(ge, ray, tip, tn, outside);
            else continue;
            if (t > 0.0f && t_min > t) { t_min = t; materialid = ge->materialid; ip = tip; normal = tn; }
        }
        if (p.nfaces && RayAABBintersect(ray, p.mesh_box)) {               // RAY_CULLING true, :258
            int best_face = -1; float bbx = 0.f, bby = 0.f; bool hit_face = false;
            if (p.use_bvh) {
                traverse_bvh(p.nodes, p.tris, ray, t_min, best_face, bbx, bby, hit_face);
            } else {
                for (int f = 0; f < p.nfaces; ++f) {
                    const ptd_face* fc = &p.faces[f];
                    float bx, by;
                    float t = triangleParam(fc->v[0], fc->v[1], fc->v[2], ray, bx, by);
                    consider_face(t, f, bx, by, t_min, best_face, bbx, bby, hit_face);
                }
            }
            if (hit_face) {
                const ptd_face* fc = &p.faces[best_face];
                materialid = fc->materialid;
                triangleFinish(fc, bbx, bby, ip, normal);
            }
        }
        float isx_t; v3 isx_n = V(0, 0, 0); int isx_mat = 0;
        if (materialid == -1) {
            isx_t = -1.0f;
        } else {
            isx_t = t_min; isx_mat = materialid; isx_n = normalize(normal);
        }
        const bool hit = isx_t >= 0;
        const int col = pixelIndex % p.W, row = pixelIndex / p.W;
        const size_t mirrored = (size_t)(p.W - col - 1) + (size_t)row * p.W;          // x-mirror of copy_data / :297-299
        if (FIRST && p.iter == 1) {                                                     // :295-304 (+ the init memset for misses)
            p.gbuf[(size_t)p.P * 3 + mirrored] = hit ? normal.x : 0.f;
            p.gbuf[(size_t)p.P * 4 + mirrored] = hit ? normal.y : 0.f;
            p.gbuf[(size_t)p.P * 5 + mirrored] = hit ? normal.z : 0.f;
            p.gbuf[(size_t)p.P * 6 + mirrored] = hit ? isx_t : 0.f;
        }
        if (p.trace_isx) {
            ptd_intersection r;
            memset(&r, 0, sizeof r);
            r.t = isx_t;
            if (hit) { r.surfaceNormal = isx_n; r.materialId = isx_mat; r.is_inside = !outside; r.intersect = ip; }
            p.trace_isx[(size_t)p.bounce * p.P + idx] = r;
        }
        if (p.sort_keys) p.sort_keys[idx] = isx_mat;                                   // key of the UN-compacted slot (:509 quirk)

        // ---- 3. shade (shadeMaterial, pathtrace.cu:333-390) -------------------------------------------------
        if (isx_t > 0.0f) {
            Rng rng = make_rng(p.iter, idx, rb);
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
        if (FIRST && p.iter == 1) {                                                     // :379-387
            p.gbuf[(size_t)p.P * 7 + mirrored] = hit ? color.x : 0.f;
            p.gbuf[(size_t)p.P * 8 + mirrored] = hit ? color.y : 0.f;
            p.gbuf[(size_t)p.P * 9 + mirrored] = hit ? color.z : 0.f;
        }
        keep = rb > 0;
        if (!keep) {
            // finalGather (:393-402) + copy_data (:81-94): every segment terminates exactly once per iteration, so its
            // throughput is accumulated and the radiance planes are emitted here instead of in two extra passes over P.
            float* img = p.image + (size_t)pixelIndex * 3;
            v3 acc = add(p.iter != 1 ? V(img[0], img[1], img[2]) : V(0.f, 0.f, 0.f), color);
            img[0] = acc.x; img[1] = acc.y; img[2] = acc.z;
            const float fi = (float)p.iter;
            p.gbuf[mirrored] = __fdiv_rn(acc.x, fi);
            p.gbuf[(size_t)p.P + mirrored] = __fdiv_rn(acc.y, fi);
            p.gbuf[(size_t)p.P * 2 + mirrored] = __fdiv_rn(acc.z, fi);
        }
    }

    // ---- 4. stable stream compaction (thrust::partition, :505) -----------------------------------------------
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int lane_rank = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp_kept[warp] = __popc(ballot);
    __syncthreads();
    int warp_off = 0, block_kept = 0;
#pragma unroll
    for (int w = 0; w < PT_BLOCK / 32; ++w) { if (w < warp) warp_off += s_warp_kept[w]; block_kept += s_warp_kept[w]; }
    if (tid == 0) {
        // decoupled look-back: state 1 = tile aggregate, 2 = inclusive prefix (value in the low 32 bits)
        int excl = 0;
        if (tile == 0) {
            st_status(&p.status[0], (2ull << 32) | (unsigned)block_kept);
        } else {
            st_status(&p.status[tile], (1ull << 32) | (unsigned)block_kept);
            int j = tile - 1;
            for (;;) {
                unsigned long long s = ld_status(&p.status[j]);
                unsigned st = (unsigned)(s >> 32);
                if (st == 0) continue;
                excl += (int)(unsigned)s;
                if (st == 2) break;
                --j;
            }
            st_status(&p.status[tile], (2ull << 32) | (unsigned)(excl + block_kept));
        }
        s_excl = excl;
        if (base + PT_BLOCK >= n) p.counts[p.bounce + 1] = excl + block_kept;            // last tile publishes the live count
    }
    const int local_rank = warp_off + lane_rank;
    if (keep) {
        float* w = reinterpret_cast<float*>(s_words) + local_rank * PT_WORDS;
        w[0] = ray.o.x; w[1] = ray.o.y; w[2] = ray.o.z; w[3] = ray.d.x; w[4] = ray.d.y; w[5] = ray.d.z;
        w[6] = color.x; w[7] = color.y; w[8] = color.z; w[9] = __int_as_float(pixelIndex); w[10] = __int_as_float(rb);
    }
    __syncthreads();
    const int excl = s_excl;
    {
        uint32_t* dst = reinterpret_cast<uint32_t*>(p.dst) + (size_t)excl * PT_WORDS;
        const int nwords = block_kept * PT_WORDS;
        for (int i = tid; i < nwords; i += PT_BLOCK) dst[i] = s_words[i];               // contiguous, 128 B per warp store
    }
    if (p.dead && active && !keep) {
        // rejected items end up behind the survivors in REVERSE order (thrust CUDA back end) and are never moved again
        const int rej_before = base - excl + (tid - local_rank);
        ptd_path_segment ps;
        ps.ray.origin = ray.o; ps.ray.direction = ray.d; ps.color = color; ps.pixelIndex = pixelIndex; ps.remainingBounces = rb;
        p.dead[(size_t)(n - 1 - rej_before)] = ps;
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
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
    int W = 0, H = 0, P = 0, depth = 0, ngeoms = 0, nmaterials = 0, nfaces = 0, ntiles = 0;
    ptd_camera cam;
    ptd_aabb mesh_box;
    ptd_geom* d_geoms = nullptr; ptd_aabb* d_geom_bounds = nullptr; ptd_material* d_materials = nullptr; ptd_face* d_faces = nullptr;
    float4* d_nodes = nullptr; float4* d_tris = nullptr;
    ptd_path_segment* d_paths[3] = {nullptr, nullptr, nullptr};
    ptd_path_segment* d_dead = nullptr;
    float* d_image = nullptr; float* d_gbuf_own = nullptr;
    unsigned char* d_ctl = nullptr; size_t ctl_bytes = 0;     // counts | tickets | status (memset once per frame)
    int* d_counts = nullptr; int* d_ticket = nullptr; unsigned long long* d_status = nullptr;
    int* d_keys = nullptr; int* d_hist = nullptr; int sort_blocks = 0;
    ptd_path_segment* d_trace_paths = nullptr; ptd_intersection* d_trace_isx = nullptr;
    int final_buf = 0;
    int launches = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> events;
    int timed_launches = 0;
    int bvh_nodes = 0, bvh_leaves = 0, bvh_max_leaf = 0, bvh_max_depth = 0;
};

extern "C" int ptd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void ptd_pt_destroy(ptd_pt* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_geoms); cudaFree(h->d_geom_bounds); cudaFree(h->d_materials); cudaFree(h->d_faces); cudaFree(h->d_nodes); cudaFree(h->d_tris);
    for (int i = 0; i < 3; ++i) cudaFree(h->d_paths[i]);
    cudaFree(h->d_dead); cudaFree(h->d_image); cudaFree(h->d_gbuf_own); cudaFree(h->d_ctl); cudaFree(h->d_keys); cudaFree(h->d_hist);
    cudaFree(h->d_trace_paths); cudaFree(h->d_trace_isx);
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    delete h;
}

extern "C" ptd_status ptd_pt_create(const ptd_scene* sc, int device, unsigned flags, ptd_pt** out) {
    if (!sc || !out) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create: null argument");
    *out = nullptr;
    if (ptd_device_count() <= device || device < 0) PTD_FAIL(PTD_ERR_CUDA, "ptd_pt_create: CUDA device %d not available (no CPU fallback exists)", device);
    if (sc->trace_depth < 1 || sc->trace_depth > 1023) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create: trace depth %d out of range", sc->trace_depth);
    for (const ptd_geom& g : sc->geoms)
        if (g.materialid < 0 || g.materialid >= (int)sc->materials.size()) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create: geom material id %d out of range", g.materialid);
    for (const ptd_face& f : sc->faces)
        if (f.materialid < 0 || f.materialid >= (int)sc->materials.size()) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_create: mesh material id %d out of range", f.materialid);
    CUDA_TRY(cudaSetDevice(device));
    ptd_pt* h = new ptd_pt();
    h->device = device; h->flags = flags;
    h->cam = sc->camera; h->mesh_box = sc->mesh_box;
    h->W = sc->camera.res_x; h->H = sc->camera.res_y; h->P = h->W * h->H; h->depth = sc->trace_depth;
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
        if (bvh.max_depth > PT_STACK) { ptd_set_error("ptd_pt_create: BVH depth %d exceeds traversal stack %d", bvh.max_depth, PT_STACK); ptd_pt_destroy(h); return PTD_ERR_UNSUPPORTED; }
        ALLOC(h->d_nodes, sizeof(PtdBvhWide) * bvh.wide.size());
        UPLOAD(h->d_nodes, bvh.wide.data(), sizeof(PtdBvhWide) * bvh.wide.size());
        ALLOC(h->d_tris, sizeof(PtdBvhTri) * bvh.tris.size());
        UPLOAD(h->d_tris, bvh.tris.data(), sizeof(PtdBvhTri) * bvh.tris.size());
    }
    const bool sort = (flags & PTD_PT_SORT_MATERIAL) != 0;
    for (int i = 0; i < (sort ? 3 : 2); ++i) ALLOC(h->d_paths[i], sizeof(ptd_path_segment) * P);
    if (flags & PTD_PT_KEEP_TERMINATED) { ALLOC(h->d_dead, sizeof(ptd_path_segment) * P); cudaMemset(h->d_dead, 0, sizeof(ptd_path_segment) * P); }
    ALLOC(h->d_image, sizeof(float) * 3 * P);
    cudaMemset(h->d_image, 0, sizeof(float) * 3 * P);
    // control block: counts[depth+1] | tickets[depth] | status[depth][ntiles]
    size_t off_counts = 0, off_ticket = ((size_t)(h->depth + 1) * 4 + 15) / 16 * 16, off_status = off_ticket + ((size_t)h->depth * 4 + 15) / 16 * 16;
    h->ctl_bytes = off_status + (size_t)h->depth * h->ntiles * 8;
    ALLOC(h->d_ctl, h->ctl_bytes);
    h->d_counts = (int*)(h->d_ctl + off_counts); h->d_ticket = (int*)(h->d_ctl + off_ticket); h->d_status = (unsigned long long*)(h->d_ctl + off_status);
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
    const size_t smem = PT_BLOCK * PT_WORDS * 4 + (sizeof(ptd_geom) + sizeof(ptd_aabb)) * 64;
    cudaFuncSetAttribute(pt_bounce<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(pt_bounce<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    CUDA_TRY(cudaDeviceSynchronize());
    *out = h;
    return PTD_OK;
}

extern "C" ptd_status ptd_pt_render(ptd_pt* h, const ptd_camera* cam, int iter, float* gbuf, void* stream_) {
    if (!h || iter < 1) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_render: bad argument");
    cudaStream_t st = (cudaStream_t)stream_;
    CUDA_TRY(cudaSetDevice(h->device));
    if (cam && (cam->res_x != h->W || cam->res_y != h->H)) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_render: camera resolution %dx%d differs from the handle's %dx%d", cam->res_x, cam->res_y, h->W, h->H);
    if (!gbuf) {
        if (!h->d_gbuf_own) { CUDA_TRY(cudaMalloc((void**)&h->d_gbuf_own, sizeof(float) * 10 * (size_t)h->P)); CUDA_TRY(cudaMemset(h->d_gbuf_own, 0, sizeof(float) * 10 * (size_t)h->P)); }
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
    p.counts = h->d_counts; p.gbuf = gbuf; p.image = h->d_image; p.dead = h->d_dead;
    p.sort_keys = h->d_keys; p.trace_paths = h->d_trace_paths; p.trace_isx = h->d_trace_isx;
    CUDA_TRY(cudaMemsetAsync(h->d_ctl, 0, h->ctl_bytes, st));
    const size_t smem = PT_BLOCK * PT_WORDS * 4 + (p.geoms_in_smem ? (sizeof(ptd_geom) + sizeof(ptd_aabb)) * h->ngeoms : 0);
    const bool sort = (h->flags & PTD_PT_SORT_MATERIAL) != 0;
    int cur = 0;
    h->launches = 0;
    int nmark = 0;
    auto mark = [&]() {
        if (!h->profiling) return;
        if ((int)h->events.size() <= nmark) { cudaEvent_t e; cudaEventCreate(&e); h->events.push_back(e); }
        cudaEventRecord(h->events[nmark++], st);
    };
    mark();
    for (int b = 0; b < h->depth; ++b) {
        const int nxt = (cur + 1) % (sort ? 3 : 2);
        p.bounce = b;
        p.src = h->d_paths[cur]; p.dst = h->d_paths[nxt];
        p.status = h->d_status + (size_t)b * h->ntiles;
        p.ticket = h->d_ticket + b;
        if (b == 0) pt_bounce<true><<<h->ntiles, PT_BLOCK, smem, st>>>(p);
        else pt_bounce<false><<<h->ntiles, PT_BLOCK, smem, st>>>(p);
        h->launches++;
        mark();
        cur = nxt;
        if (sort && b + 1 < h->depth) {
            const int nb = std::max(h->nmaterials, 1), srt = (cur + 1) % 3;
            sort_hist<<<h->sort_blocks, SORT_TILE, nb * sizeof(int), st>>>(h->d_keys, h->d_counts + b + 1, nb, h->sort_blocks, h->d_hist);
            sort_scan<<<1, 1024, 0, st>>>(h->d_hist, nb * h->sort_blocks);
            sort_scatter<<<h->sort_blocks, SORT_TILE, 0, st>>>(h->d_keys, h->d_counts + b + 1, nb, h->sort_blocks, h->d_hist, h->d_paths[cur], h->d_paths[srt]);
            h->launches += 3;
            cur = srt;
        }
    }
    h->final_buf = cur;
    if (h->profiling) h->timed_launches = nmark - 1;
    CUDA_TRY(cudaGetLastError());
    return PTD_OK;
}

extern "C" ptd_status ptd_pt_render_host(ptd_pt* h, const ptd_camera* cam, int iter, float* host_tensor) {
    if (!h || !host_tensor) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_render_host: null argument");
    ptd_status rc = ptd_pt_render(h, cam, iter, nullptr, nullptr);
    if (rc != PTD_OK) return rc;
    CUDA_TRY(cudaMemcpy(host_tensor, h->d_gbuf_own, sizeof(float) * 10 * (size_t)h->P, cudaMemcpyDeviceToHost));   // pathtrace.cu:525
    return PTD_OK;
}

extern "C" ptd_status ptd_pt_export_rgba8(ptd_pt* h, int iter, unsigned char* pbo, void* stream_) {
    if (!h || !pbo || iter < 1) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_export_rgba8: bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    dim3 b(8, 8), g((h->W + 7) / 8, (h->H + 7) / 8);
    export_rgba8<<<g, b, 0, (cudaStream_t)stream_>>>(h->d_image, h->W, h->H, iter, (uchar4*)pbo);
    CUDA_TRY(cudaGetLastError());
    return PTD_OK;
}

extern "C" ptd_status ptd_pt_live_counts(ptd_pt* h, int* counts, int capacity, int* bounces_run) {
    if (!h || !counts || capacity < h->depth) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_live_counts: need capacity >= trace depth %d", h ? h->depth : 0);
    CUDA_TRY(cudaSetDevice(h->device));
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
    CUDA_TRY(cudaMemcpy(host, h->d_dead, sizeof(ptd_path_segment) * (size_t)h->P, cudaMemcpyDeviceToHost));
    return PTD_OK;
}
extern "C" ptd_status ptd_pt_dump_image(ptd_pt* h, float* host_rgb) {
    if (!h || !host_rgb) PTD_FAIL(PTD_ERR_ARG, "ptd_pt_dump_image: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
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
