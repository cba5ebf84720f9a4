// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Headless driver around the UNMODIFIED reference path tracer.  It is compiled, by
// oracle/build_ref.sh, as ONE translation unit together with the reference's own
// Inference/src/pathtrace.cu (read from /root/reference where it lies; nothing is copied
// into this repo).  The result, oracle/_ref/libref_pt*.so, gives the tests
//
//   * oracle A  - ref_gpu_render(): the reference's pathtraceInit()/pathtrace() run verbatim on
//     the GPU (pathtrace.cu:96-129, :422-528), with a per-bounce dump of dev_paths and
//     dev_intersections taken at the reference's own cudaDeviceSynchronize() (pathtrace.cu:483);
//   * oracle B  - ref_cpu_render(): the same five kernel bodies (pathtrace.cu:155-182, 200-306,
//     333-390, 393-402, 81-94) restated as host loops that call the reference's own
//     __host__ __device__ functions (intersections.h, interactions.h) compiled as host code.
//     It runs without a GPU, pins oracle/pt_oracle.c and is the timed CPU baseline
//     (cpu_baseline.kind == "reference").
//
// Decisions frozen here: D4 dev_paths zero-initialised before ray
// generation, D5 non-square frames allowed (the assert at pathtrace.cu:426 is compiled out
// with -DNDEBUG), iter == 1 in the frame loop, reference default macros.
#include <thrust/partition.h>
#include <thrust/sort.h>
#include <cfloat>
#include <cstring>
#include <vector>
#include <algorithm>
#include <chrono>
#include <cuda_runtime.h>
#include "sceneStructs.h"

struct RefTrace {
    bool enabled = false;
    std::vector<int> n;                                  // live paths entering bounce d
    std::vector<std::vector<PathSegment>> paths;         // dev_paths[0..n) entering bounce d
    std::vector<std::vector<ShadeableIntersection>> isx; // dev_intersections[0..n) of bounce d
};
static RefTrace g_trace;

// Called through the macro below at the reference's per-bounce sync point, i.e. after
// computeIntersections(depth) and before shadeMaterial(depth).
static cudaError_t ref_hook_bounce(int depth, int num_paths, const PathSegment* d_paths,
                                   const ShadeableIntersection* d_isx) {
    cudaError_t e = cudaDeviceSynchronize();
    if (g_trace.enabled) {
        g_trace.n.push_back(num_paths);
        g_trace.paths.emplace_back(num_paths);
        g_trace.isx.emplace_back(num_paths);
        cudaMemcpy(g_trace.paths.back().data(), d_paths, sizeof(PathSegment) * (size_t)num_paths, cudaMemcpyDeviceToHost);
        cudaMemcpy(g_trace.isx.back().data(), d_isx, sizeof(ShadeableIntersection) * (size_t)num_paths, cudaMemcpyDeviceToHost);
    }
    (void)depth;
    return e;
}

#define cudaDeviceSynchronize() ref_hook_bounce(depth, num_paths, dev_paths, dev_intersections)
#include "pathtrace.cu"   // the reference file itself, found through -I/root/reference/Inference/src
#undef cudaDeviceSynchronize

// ------------------------------------------------------------------------------------------------
// oracle B: host loops around the reference's own host/device functions
// ------------------------------------------------------------------------------------------------
static void cpu_raygen(const Camera& cam, int iter, int traceDepth, PathSegment* paths) {
    const int W = cam.resolution.x, H = cam.resolution.y;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int index = x + y * W;
            thrust::default_random_engine rng = makeSeededRandomEngine(iter, index, paths[index].remainingBounces);
            thrust::uniform_real_distribution<float> u01(-0.5, 0.5);
            PathSegment& segment = paths[index];
            segment.ray.origin = cam.position;
            segment.color = glm::vec3(1.0f, 1.0f, 1.0f);
            // The device build draws the x jitter first (hard part 1); sequence it.
            float jx = u01(rng);
            float jy = u01(rng);
            segment.ray.direction = glm::normalize(cam.view
                - cam.right * cam.pixelLength.x * ((float)x - (float)cam.resolution.x * 0.5f + jx)
                - cam.up * cam.pixelLength.y * ((float)y - (float)cam.resolution.y * 0.5f + jy));
            segment.pixelIndex = index;
            segment.remainingBounces = traceDepth;
        }
}

static void cpu_intersect(int depth, int num_paths, const PathSegment* paths, Scene* sc, float* tensor,
                          ShadeableIntersection* isx, int iter, int width) {
    const int geoms_size = (int)sc->geoms.size();
    const int face_size = (int)sc->faces.size();
#pragma omp parallel for schedule(dynamic, 256)
    for (int path_index = 0; path_index < num_paths; ++path_index) {
        PathSegment pathSegment = paths[path_index];
        float t;
        glm::vec3 intersect_point, normal, tmp_intersect, tmp_normal;
        float t_min = FLT_MAX;
        int materialid = -1;
        bool outside = true;
        for (int i = 0; i < geoms_size; i++) {
            Geom& geom = sc->geoms[i];
            if (geom.type == CUBE) t = boxIntersectionTest(geom, pathSegment.ray, tmp_intersect, tmp_normal, outside);
            else if (geom.type == SPHERE) t = sphereIntersectionTest(geom, pathSegment.ray, tmp_intersect, tmp_normal, outside);
            if (t > 0.0f && t_min > t) { t_min = t; materialid = geom.materialid; intersect_point = tmp_intersect; normal = tmp_normal; }
        }
        if (face_size && RayAABBintersect(pathSegment.ray, sc->mesh_box)) {
            for (int i = 0; i < face_size; i++) {
                t = triangleIntersectionTest(sc->faces[i], pathSegment.ray, tmp_intersect, tmp_normal, outside);
                if (t > 0.0f && t_min > t) { t_min = t; materialid = sc->faces[i].materialid; intersect_point = tmp_intersect; normal = tmp_normal; }
            }
        }
        if (materialid == -1) {
            isx[path_index].t = -1.0f;
        } else {
            isx[path_index].t = t_min;
            isx[path_index].materialId = materialid;
            isx[path_index].surfaceNormal = glm::normalize(normal);
            isx[path_index].is_inside = !outside;
            isx[path_index].intersect = intersect_point;
        }
        if (depth == 0 && iter == 1 && isx[path_index].t >= 0) {
            int y = path_index % width, x = path_index / width;
            int new_1d = (width - y - 1) + x * width;
            tensor[(size_t)num_paths * 3 + new_1d] = normal.x;
            tensor[(size_t)num_paths * 4 + new_1d] = normal.y;
            tensor[(size_t)num_paths * 5 + new_1d] = normal.z;
            tensor[(size_t)num_paths * 6 + new_1d] = isx[path_index].t;
        }
    }
}

static void cpu_shade(int iter, int num_paths, const ShadeableIntersection* isx, PathSegment* paths,
                      Scene* sc, float* tensor, int depth, int width) {
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < num_paths; ++idx) {
        if (paths[idx].remainingBounces == 0) continue;
        ShadeableIntersection intersection = isx[idx];
        if (intersection.t > 0.0f) {
            thrust::default_random_engine rng = makeSeededRandomEngine(iter, idx, paths[idx].remainingBounces);
            Material material = sc->materials[intersection.materialId];
            glm::vec3 materialColor = material.color;
            if (material.emittance > 0.0f) {
                paths[idx].remainingBounces = 0;
                paths[idx].color = paths[idx].color * materialColor * material.emittance;
            } else {
                scatterRay(paths[idx], intersection, material, rng);
                --paths[idx].remainingBounces;
            }
        } else {
            paths[idx].color = glm::vec3(0.0f);
            paths[idx].remainingBounces = 0;
        }
        if (depth == 0 && iter == 1 && intersection.t >= 0) {
            int y = idx % width, x = idx / width;
            int new_1d = (width - y - 1) + x * width;
            tensor[(size_t)num_paths * 7 + new_1d] = paths[idx].color.x;
            tensor[(size_t)num_paths * 8 + new_1d] = paths[idx].color.y;
            tensor[(size_t)num_paths * 9 + new_1d] = paths[idx].color.z;
        }
    }
}

// thrust::partition on the CUDA back end: selected items keep their order, rejected items follow in
// REVERSE order (thrust/system/cuda/detail/partition.h).
static int cpu_partition(PathSegment* paths, int n) {
    PathSegment* mid = std::stable_partition(paths, paths + n, [](const PathSegment& p) { return p.remainingBounces > 0; });
    std::reverse(mid, paths + n);
    return (int)(mid - paths);
}

extern "C" {

int ref_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(PathSegment);
        case 1: return (int)sizeof(ShadeableIntersection);
        case 2: return (int)sizeof(Geom);
        case 3: return (int)sizeof(Face);
        case 4: return (int)sizeof(Material);
        case 5: return (int)sizeof(Camera);
        case 6: return (int)sizeof(MeshBoundingBox);
    }
    return -1;
}

int ref_flags(void) {   // bit0 compaction, bit1 material sort, bit2 aabb cull, bit3 AA
    return (STREAM_COMPACTION ? 1 : 0) | (SORT_MATERIAL ? 2 : 0) | (RAY_CULLING ? 4 : 0) | (AA ? 8 : 0);
}

void* ref_scene_load(const char* path) {
    Scene* s = nullptr;
    try { s = new Scene(std::string(path)); } catch (...) { return nullptr; }
    return s;
}
int ref_scene_counts(void* sp, int* out /*[5]: geoms, materials, faces, depth, iterations*/) {
    Scene* s = (Scene*)sp;
    out[0] = (int)s->geoms.size(); out[1] = (int)s->materials.size(); out[2] = (int)s->faces.size();
    out[3] = s->state.traceDepth; out[4] = (int)s->state.iterations;
    return 0;
}
const void* ref_scene_geoms(void* sp) { return ((Scene*)sp)->geoms.data(); }
const void* ref_scene_materials(void* sp) { return ((Scene*)sp)->materials.data(); }
const void* ref_scene_faces(void* sp) { return ((Scene*)sp)->faces.data(); }
const void* ref_scene_meshbox(void* sp) { return &((Scene*)sp)->mesh_box; }
void* ref_scene_camera(void* sp) { return &((Scene*)sp)->state.camera; }
void ref_scene_set_camera(void* sp, const void* cam) { memcpy(&((Scene*)sp)->state.camera, cam, sizeof(Camera)); }
void ref_scene_set_depth(void* sp, int d) { ((Scene*)sp)->state.traceDepth = d; }

void ref_trace_enable(int on) { g_trace.enabled = on != 0; }
void ref_trace_clear(void) { g_trace.n.clear(); g_trace.paths.clear(); g_trace.isx.clear(); }
int ref_trace_bounces(void) { return (int)g_trace.n.size(); }
int ref_trace_count(int b) { return g_trace.n[b]; }
const void* ref_trace_paths(int b) { return g_trace.paths[b].data(); }
const void* ref_trace_isx(int b) { return g_trace.isx[b].data(); }

// oracle A.  Returns 0 on success, else the CUDA error code.  host_tensor is float[10*P];
// image (float[3*P]) and final_paths (PathSegment[P]) may be NULL.
int ref_gpu_render(void* sp, int iter, float* host_tensor, float* image, void* final_paths, float* ms) {
    Scene* s = (Scene*)sp;
    const Camera& cam = s->state.camera;
    const int P = cam.resolution.x * cam.resolution.y;
    float* saved = s->state.host_tensor;
    s->state.host_tensor = host_tensor;
    ref_trace_clear();
    pathtraceFree();                       // main.cpp:143-146 does Free+Init before every frame
    pathtraceInit(s);
    cudaMemset(dev_paths, 0, sizeof(PathSegment) * (size_t)P);   // decision D4
    uchar4* pbo = nullptr;
    cudaMalloc(&pbo, sizeof(uchar4) * (size_t)P);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    pathtrace(pbo, 0, iter);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (ms) cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (image) cudaMemcpy(image, dev_image, sizeof(glm::vec3) * (size_t)P, cudaMemcpyDeviceToHost);
    if (final_paths) cudaMemcpy(final_paths, dev_paths, sizeof(PathSegment) * (size_t)P, cudaMemcpyDeviceToHost);
    cudaFree(pbo);
    s->state.host_tensor = saved;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = ::cudaDeviceSynchronize();
    return (int)e;
}

// oracle B.  sort_material mirrors SORT_MATERIAL (keys = materialId of the un-compacted
// intersection slots [0,n), stable merge sort; pathtrace.cu:508-510).  Returns sum of live paths.
long long ref_cpu_render(void* sp, int iter, int sort_material, float* host_tensor, float* image,
                         void* final_paths, double* ms) {
    Scene* s = (Scene*)sp;
    const Camera cam = s->state.camera;
    const int W = cam.resolution.x, H = cam.resolution.y, P = W * H;
    const int traceDepth = s->state.traceDepth;
    std::vector<PathSegment> paths(P);
    memset((void*)paths.data(), 0, sizeof(PathSegment) * (size_t)P);          // decision D4
    std::vector<ShadeableIntersection> isx(P);
    std::vector<glm::vec3> img(P, glm::vec3(0.0f));                    // pathtrace.cu:102
    memset(host_tensor, 0, sizeof(float) * 10 * (size_t)P);            // pathtrace.cu:119
    ref_trace_clear();
    auto t0 = std::chrono::high_resolution_clock::now();
    cpu_raygen(cam, iter, traceDepth, paths.data());
    int depth = 0, num_paths = P;
    long long sum = 0;
    bool done = false;
    while (!done) {
        memset((void*)isx.data(), 0, sizeof(ShadeableIntersection) * (size_t)P);   // pathtrace.cu:478
        cpu_intersect(depth, num_paths, paths.data(), s, host_tensor, isx.data(), iter, W);
        if (g_trace.enabled) {
            g_trace.n.push_back(num_paths);
            g_trace.paths.emplace_back(paths.begin(), paths.begin() + num_paths);
            g_trace.isx.emplace_back(isx.begin(), isx.begin() + num_paths);
        }
        sum += num_paths;
        cpu_shade(iter, num_paths, isx.data(), paths.data(), s, host_tensor, depth, W);
        depth++;
        num_paths = cpu_partition(paths.data(), num_paths);
        if (sort_material) {
            std::vector<int> order(num_paths);
            for (int i = 0; i < num_paths; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return isx[a].materialId < isx[b].materialId; });
            std::vector<PathSegment> tmp(num_paths);
            for (int i = 0; i < num_paths; ++i) tmp[i] = paths[order[i]];
            std::copy(tmp.begin(), tmp.end(), paths.begin());
        }
        done = (num_paths == 0 || depth == traceDepth);
    }
    for (int i = 0; i < P; ++i) img[paths[i].pixelIndex] += paths[i].color;     // finalGather
    for (int y = 0; y < H; ++y)                                                  // copy_data
        for (int x = 0; x < W; ++x) {
            int src = (W - x - 1) + y * W, dst = x + y * W;
            glm::vec3 pix = img[src];
            host_tensor[dst] = pix.x / (float)iter;
            host_tensor[dst + P] = pix.y / (float)iter;
            host_tensor[dst + 2 * (size_t)P] = pix.z / (float)iter;
        }
    auto t1 = std::chrono::high_resolution_clock::now();
    if (ms) *ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (image) memcpy(image, img.data(), sizeof(glm::vec3) * (size_t)P);
    if (final_paths) memcpy(final_paths, paths.data(), sizeof(PathSegment) * (size_t)P);
    return sum;
}

// Bounded CPU-baseline sample for mesh scenes (brute force is O(F) per ray): first bounce of every
// `row_stride`-th image row only.  Returns rays traced; *ms is the wall time.
long long ref_cpu_first_bounce_rows(void* sp, int iter, int row_stride, double* ms) {
    Scene* s = (Scene*)sp;
    const Camera cam = s->state.camera;
    const int W = cam.resolution.x, H = cam.resolution.y, P = W * H;
    std::vector<PathSegment> paths(P);
    memset((void*)paths.data(), 0, sizeof(PathSegment) * (size_t)P);
    cpu_raygen(cam, iter, s->state.traceDepth, paths.data());
    std::vector<PathSegment> sel;
    for (int y = 0; y < H; y += row_stride) sel.insert(sel.end(), paths.begin() + (size_t)y * W, paths.begin() + (size_t)(y + 1) * W);
    std::vector<ShadeableIntersection> isx(sel.size());
    std::vector<float> scratch(10 * (size_t)sel.size());
    auto t0 = std::chrono::high_resolution_clock::now();
    cpu_intersect(1 /* no g-buffer write */, (int)sel.size(), sel.data(), s, scratch.data(), isx.data(), iter, W);
    auto t1 = std::chrono::high_resolution_clock::now();
    if (ms) *ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    return (long long)sel.size();
}

}  // extern "C"
