/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C CPU restatement of the reference's 1-spp path-trace iteration (hot path HP-1):
 * Inference/src/pathtrace.cu, intersections.h, interactions.h and the camera-orbit part of
 * main.cpp, with GLM 0.9.6.3's expression trees and thrust::minstd_rand written out as scalar
 * fp32 arithmetic.  Every function cites the reference lines it follows.
 *
 * Pinned (tests/test_oracle_pt.py) against oracle/_ref's ref_cpu_render() -- the reference's own
 * __host__ __device__ functions compiled as host code -- bit for bit on the committed scenes, and
 * against the golden vectors under tests/golden/ produced by tests/tools/make_golden_pt.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC pt_oracle.c -lm   (no FMA contraction:
 * x86-64 host code of the reference is compiled without FMA as well).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- record layouts, Inference/src/sceneStructs.h ---------------------------------------------- */
typedef struct { float x, y, z; } v3;
typedef struct { v3 origin, direction; } Ray;                                           /* :15-18 */
typedef struct { int type, materialid; v3 translation, rotation, scale;
                 float transform[16], inverseTransform[16], invTranspose[16]; v3 vel; } Geom;      /* :20-30, 248 B */
typedef struct { v3 v[3]; v3 n[3]; int materialid; } Face;                              /* :40-44, 76 B */
typedef struct { v3 color; float specex; v3 speccolor; float hasReflective, hasRefractive,
                 indexOfRefraction, emittance; } Material;                              /* :46-56, 44 B */
typedef struct { int resx, resy; v3 position, lookAt, view, up, right; float fovx, fovy, plx, ply; } Camera; /* :58-67, 84 B */
typedef struct { Ray ray; v3 color; int pixelIndex, remainingBounces; } PathSegment;   /* :77-82, 44 B */
typedef struct { v3 lb, ub; } MeshBox;                                                  /* :84-87 */
typedef struct { float t; v3 surfaceNormal; int materialId; unsigned char is_inside, pad[3]; v3 intersect; } Isect; /* :91-97, 36 B */
enum { SPHERE = 0, CUBE = 1 };                                                          /* :10-13 */

typedef char check_sizes[(sizeof(Geom) == 248 && sizeof(Face) == 76 && sizeof(Material) == 44 && sizeof(Camera) == 84 &&
                          sizeof(PathSegment) == 44 && sizeof(Isect) == 36) ? 1 : -1];

#define PI_F 3.1415926535897932384626422832795028841971f          /* utilities.h:13 */
#define TWO_PI_F 6.2831853071795864769252867665590057683943f      /* utilities.h:14 */
#define SQRT_OF_ONE_THIRD_F 0.5773502691896257645091487805019574556476f

/* ---- GLM 0.9.6.3 expression trees (external/include/glm/detail/func_geometric.inl) -------------- */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { v3 t = mulv(a, b); return t.x + t.y + t.z; }                 /* :66-73 */
static inline v3 cross(v3 x, v3 y) {                                                            /* :134-143 */
    return V(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
static inline v3 normalize(v3 a) { return muls(a, 1.0f / sqrtf(dot(a, a))); }  /* :153-159 + func_exponential.inl:148-153 */
static inline float length(v3 a) { return sqrtf(dot(a, a)); }                                    /* :95-101 */
static inline v3 reflect(v3 I, v3 N) { return sub(I, muls(muls(N, dot(N, I)), 2.0f)); }          /* :174-179 */
static inline v3 glm_refract(v3 I, v3 N, float eta) {                                            /* :192-200 */
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    v3 r = sub(muls(I, eta), muls(N, eta * d + sqrtf(k)));
    return muls(r, (float)(k >= 0.0f));
}
/* mat4 * vec4 -> xyz, detail/type_mat4x4.inl:617-628: (m0*v0 + m1*v1) + (m2*v2 + m3*v3); column major */
static inline v3 mulMV(const float* m, v3 v, float w) {
    v3 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * w);
    return r;
}
static inline float glm_min(float x, float y) { return x < y ? x : y; }     /* func_common.inl:409-414 */
static inline float glm_max(float x, float y) { return x > y ? x : y; }     /* func_common.inl:430-435 */

/* ---- RNG: intersections.h:12-20, pathtrace.cu:52-56, thrust::minstd_rand ------------------------ */
static inline uint32_t utilhash(uint32_t a) {
    a = (a + 0x7ed55d16) + (a << 12);
    a = (a ^ 0xc761c23c) ^ (a >> 19);
    a = (a + 0x165667b1) + (a << 5);
    a = (a + 0xd3a2646c) ^ (a << 9);
    a = (a + 0xfd7046c5) + (a << 3);
    a = (a ^ 0xb55a4f09) ^ (a >> 16);
    return a;
}
typedef struct { uint32_t x; } Rng;
static inline Rng make_rng(int iter, int index, int depth) {
    uint32_t h = utilhash((1u << 31) | ((uint32_t)depth << 22) | (uint32_t)iter) ^ utilhash((uint32_t)index);
    Rng r;                                   /* linear_congruential_engine.inl:45-56: seed(s) */
    r.x = h % 2147483647u;
    if (r.x == 0) r.x = 1;
    return r;
}
static inline uint32_t rng_next(Rng* r) {    /* thrust/random/detail/mod.h: Schrage, a=48271, m=2^31-1 */
    const uint32_t a = 48271u, m = 2147483647u, q = m / a, rr = m % a;
    uint32_t x = r->x;
    uint32_t t1 = a * (x % q), t2 = rr * (x / q);
    x = (t1 >= t2) ? (t1 - t2) : (m - t2 + t1);
    r->x = x;
    return x;
}
static inline float rng_uniform(Rng* r, float lo, float hi) {   /* uniform_real_distribution.inl:60-73 */
    float result = (float)(rng_next(r) - 1u);
    result /= (1.0f + (float)(2147483646u - 1u));
    return (result * (hi - lo)) + lo;
}

/* ---- intersections.h ---------------------------------------------------------------------------- */
static inline v3 getPointOnRay(Ray r, float t) {                                       /* :27-29 */
    return add(r.origin, muls(normalize(r.direction), (t - .0001f)));
}

static float boxIntersectionTest(const Geom* box, Ray r, v3* ip, v3* normal, int* outside) {   /* :52-94 */
    Ray q;
    q.origin = mulMV(box->inverseTransform, r.origin, 1.0f);
    q.direction = normalize(mulMV(box->inverseTransform, r.direction, 0.0f));
    float tmin = -1e38f, tmax = 1e38f;
    v3 tmin_n = V(0, 0, 0), tmax_n = V(0, 0, 0);
    const float* qo = &q.origin.x;
    const float* qd = &q.direction.x;
    for (int xyz = 0; xyz < 3; ++xyz) {
        float qdxyz = qd[xyz];
        float t1 = (-0.5f - qo[xyz]) / qdxyz;
        float t2 = (+0.5f - qo[xyz]) / qdxyz;
        float ta = glm_min(t1, t2);
        float tb = glm_max(t1, t2);
        v3 n = V(0, 0, 0);
        (&n.x)[xyz] = t2 < t1 ? +1.0f : -1.0f;
        if (ta > 0 && ta > tmin) { tmin = ta; tmin_n = n; }
        if (tb < tmax) { tmax = tb; tmax_n = n; }
    }
    if (tmax >= tmin && tmax > 0) {
        *outside = 1;
        if (tmin <= 0) { tmin = tmax; tmin_n = tmax_n; *outside = 0; }
        *ip = mulMV(box->transform, getPointOnRay(q, tmin), 1.0f);
        *normal = normalize(mulMV(box->transform, tmin_n, 0.0f));
        return length(sub(r.origin, *ip));
    }
    return -1;
}

static float sphereIntersectionTest(const Geom* sphere, Ray r, v3* ip, v3* normal, int* outside) {  /* :106-148 */
    Ray rt;
    rt.origin = mulMV(sphere->inverseTransform, r.origin, 1.0f);
    rt.direction = normalize(mulMV(sphere->inverseTransform, r.direction, 0.0f));
    float vDotDirection = dot(rt.origin, rt.direction);
    float radicand = vDotDirection * vDotDirection - (dot(rt.origin, rt.origin) - 0.25f /* powf(.5,2) */);
    if (radicand < 0) return -1;
    float squareRoot = sqrtf(radicand);
    float firstTerm = -vDotDirection;
    float t1 = firstTerm + squareRoot;
    float t2 = firstTerm - squareRoot;
    float t = 0;
    if (t1 < 0 && t2 < 0) return -1;
    else if (t1 > 0 && t2 > 0) { t = fminf(t1, t2); *outside = 1; }
    else { t = fmaxf(t1, t2); *outside = 0; }
    v3 obj = getPointOnRay(rt, t);
    *ip = mulMV(sphere->transform, obj, 1.0f);
    *normal = normalize(mulMV(sphere->invTranspose, obj, 0.0f));
    if (!*outside) *normal = neg(*normal);
    return length(sub(r.origin, *ip));
}

/* glm/gtx/intersect.inl:37-74 (back-face culling Moeller-Trumbore), then intersections.h:159-172 */
static float triangleIntersectionTest(const Face* f, Ray r, v3* ip, v3* normal) {
    v3 e1 = sub(f->v[1], f->v[0]);
    v3 e2 = sub(f->v[2], f->v[0]);
    v3 p = cross(r.direction, e2);
    float a = dot(e1, p);
    if (a < FLT_EPSILON) return -1;
    float ff = 1.0f / a;
    v3 s = sub(r.origin, f->v[0]);
    float bx = ff * dot(s, p);
    if (bx < 0.0f) return -1;
    if (bx > 1.0f) return -1;
    v3 q = cross(s, e1);
    float by = ff * dot(r.direction, q);
    if (by < 0.0f) return -1;
    if (by + bx > 1.0f) return -1;
    float bz = ff * dot(e2, q);
    if (!(bz >= 0.0f)) return -1;
    /* the reference maps (x, y, 1-x-y) onto (v0, v1, v2) for the point (sic) and (n0,n1,n2) <- (1-x-y, x, y) */
    *ip = add(add(muls(f->v[0], bx), muls(f->v[1], by)), muls(f->v[2], (1 - bx - by)));
    *normal = normalize(add(add(muls(f->n[0], (1 - bx - by)), muls(f->n[1], bx)), muls(f->n[2], by)));
    return bz;
}

static int RayAABBintersect(const Ray* ray, const MeshBox* b) {                        /* :175-200 */
    float dx = 1.0f / ray->direction.x, dy = 1.0f / ray->direction.y, dz = 1.0f / ray->direction.z;
    float t1 = (b->lb.x - ray->origin.x) * dx, t2 = (b->ub.x - ray->origin.x) * dx;
    float t3 = (b->lb.y - ray->origin.y) * dy, t4 = (b->ub.y - ray->origin.y) * dy;
    float t5 = (b->lb.z - ray->origin.z) * dz, t6 = (b->ub.z - ray->origin.z) * dz;
    float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    if (tmax < 0) return 0;
    if (tmin > tmax) return 0;
    return 1;
}

/* ---- interactions.h ------------------------------------------------------------------------------ */
static v3 calculateRandomDirectionInHemisphere(v3 normal, Rng* rng) {                   /* :13-44 */
    float up = sqrtf(rng_uniform(rng, 0, 1));
    float over = sqrtf(1 - up * up);
    float around = rng_uniform(rng, 0, 1) * TWO_PI_F;
    v3 directionNotNormal;
    if (fabsf(normal.x) < SQRT_OF_ONE_THIRD_F) directionNotNormal = V(1, 0, 0);
    else if (fabsf(normal.y) < SQRT_OF_ONE_THIRD_F) directionNotNormal = V(0, 1, 0);
    else directionNotNormal = V(0, 0, 1);
    v3 p1 = normalize(cross(normal, directionNotNormal));
    v3 p2 = normalize(cross(normal, p1));
    return add(add(muls(normal, up), muls(p1, cosf(around) * over)), muls(p2, sinf(around) * over));
}

static int refract_hw(v3 v, v3 n, float ni_over_nt, v3* refracted) {                    /* :74-85 */
    v3 uv = normalize(v);
    float dt = dot(uv, n);
    float discriminat = (float)(1.0 - (double)(ni_over_nt * ni_over_nt * (1 - dt * dt)));   /* `1.0` is a double literal */
    if (discriminat > 0) {
        *refracted = sub(muls(sub(uv, muls(n, dt)), ni_over_nt), muls(n, sqrtf(discriminat)));
        return 1;
    }
    return 0;
}

static float schlick(float cosine, float ref_idx) {                                     /* :116-120 */
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    /* pow(float, int) promotes to double pow() in C++11, host and device alike (probe in DESIGN.md) */
    return (float)((double)r0 + (double)(1 - r0) * pow((double)(1 - cosine), 5.0));
}

/* live branch of scatterRay with the reference's default macros (DIELECTRIC false, FRESNELS true,
 * MESH_NORMAL_VIEW false): interactions.h:170-259, lines :194-258 */
static void scatterRay(PathSegment* ps, const Isect* isx, const Material* m, Rng* rng) {
    v3 dir = ps->ray.direction;
    v3 color = V(1.0f, 1.0f, 1.0f);
    float reflective_prob = m->hasReflective;
    if (reflective_prob != 0 || m->hasRefractive != 0) {
        float pdf = rng_uniform(rng, 0, 1), refrac_index_ratio, cosine;
        v3 normal;
        cosine = dot(normalize(dir), isx->surfaceNormal);
        if (cosine <= 0) {
            normal = isx->surfaceNormal;
            refrac_index_ratio = 1 / m->indexOfRefraction;
            cosine = -cosine;
        } else {
            normal = neg(isx->surfaceNormal);
            refrac_index_ratio = m->indexOfRefraction;
        }
        if (refract_hw(ps->ray.direction, normal, refrac_index_ratio, &dir))
            reflective_prob = schlick(cosine, refrac_index_ratio);
        else
            reflective_prob = 1.0f;
        if (pdf < reflective_prob) {
            dir = normalize(reflect(dir, isx->surfaceNormal));
            color = m->speccolor;
        } else {
            dir = normalize(glm_refract(ps->ray.direction, normal, refrac_index_ratio));
            if (!length(dir)) {
                dir = normalize(reflect(dir, isx->surfaceNormal));
                color = m->speccolor;
            } else {
                color = m->color;
            }
        }
    } else {
        dir = normalize(calculateRandomDirectionInHemisphere(isx->surfaceNormal, rng));
        color = m->color;
    }
    ps->ray.direction = dir;
    ps->ray.origin = add(isx->intersect, muls(dir, 0.01f));
    ps->color = mulv(ps->color, color);
}

/* ---- camera: scene.cpp:142-152 and main.cpp:66-78, :126-138 -------------------------------------- */
void pto_camera_from_scene(Camera* cam, float fovy_deg) {      /* loadCamera's derived fields */
    float yscaled = tanf(fovy_deg * (PI_F / 180));
    float xscaled = (yscaled * cam->resx) / cam->resy;
    float fovx = (atanf(xscaled) * 180) / PI_F;
    cam->fovx = fovx; cam->fovy = fovy_deg;
    cam->plx = 2 * xscaled / (float)cam->resx;
    cam->ply = 2 * yscaled / (float)cam->resy;
    cam->view = normalize(sub(cam->lookAt, cam->position));
}
void pto_camera_orbit_params(const Camera* cam, float* zoom, float* phi, float* theta) {   /* main.cpp:66-78 */
    v3 view = cam->view;
    v3 viewXZ = V(view.x, 0.0f, view.z), viewZY = V(0.0f, view.y, view.z);
    *phi = acosf(dot(normalize(viewXZ), V(0, 0, -1)));
    *theta = acosf(dot(normalize(viewZY), V(0, 1, 0)));
    *zoom = length(sub(cam->position, cam->lookAt));
}
void pto_camera_orbit(Camera* cam, float zoom, float phi, float theta) {                    /* main.cpp:126-138 */
    v3 cp;
    cp.x = zoom * sinf(phi) * sinf(theta);
    cp.y = zoom * cosf(theta);
    cp.z = zoom * cosf(phi) * sinf(theta);
    cam->view = neg(normalize(cp));
    v3 v = cam->view, u = V(0, 1, 0);
    v3 r = cross(v, u);
    cam->up = cross(r, v);
    cam->right = r;                         /* not normalised in the reference */
    cam->position = add(cp, cam->lookAt);
}

/* ---- the five kernel bodies ----------------------------------------------------------------------- */
typedef struct {
    int ngeoms, nmaterials, nfaces, trace_depth;
    const Geom* geoms; const Material* materials; const Face* faces; const MeshBox* mesh_box;
} SceneView;

static void raygen(const Camera* cam, int iter, int traceDepth, PathSegment* paths) {    /* pathtrace.cu:155-182 */
    const int W = cam->resx, H = cam->resy;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int index = x + y * W;
            Rng rng = make_rng(iter, index, paths[index].remainingBounces);   /* stale value: decision D4 -> 0 */
            PathSegment* s = &paths[index];
            s->ray.origin = cam->position;
            s->color = V(1.0f, 1.0f, 1.0f);
            float jx = rng_uniform(&rng, -0.5f, 0.5f);     /* device order: x jitter first */
            float jy = rng_uniform(&rng, -0.5f, 0.5f);
            v3 a = muls(muls(cam->right, cam->plx), ((float)x - (float)W * 0.5f + jx));
            v3 b = muls(muls(cam->up, cam->ply), ((float)y - (float)H * 0.5f + jy));
            s->ray.direction = normalize(sub(sub(cam->view, a), b));
            s->pixelIndex = index;
            s->remainingBounces = traceDepth;
        }
}

static void intersect(int depth, int n, const PathSegment* paths, const SceneView* sc, float* tensor,
                      Isect* isx, int iter, int width) {                                  /* pathtrace.cu:200-306 */
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        PathSegment ps = paths[i];
        float t = 0, t_min = FLT_MAX;
        v3 ip = V(0, 0, 0), normal = V(0, 0, 0), tip = V(0, 0, 0), tn = V(0, 0, 0);
        int materialid = -1, outside = 1;
        for (int g = 0; g < sc->ngeoms; ++g) {
            const Geom* ge = &sc->geoms[g];
            if (ge->type == CUBE) t = boxIntersectionTest(ge, ps.ray, &tip, &tn, &outside);
            else if (ge->type == SPHERE) t = sphereIntersectionTest(ge, ps.ray, &tip, &tn, &outside);
            if (t > 0.0f && t_min > t) { t_min = t; materialid = ge->materialid; ip = tip; normal = tn; }
        }
        if (sc->nfaces && RayAABBintersect(&ps.ray, sc->mesh_box)) {                      /* RAY_CULLING true */
            for (int f = 0; f < sc->nfaces; ++f) {
                t = triangleIntersectionTest(&sc->faces[f], ps.ray, &tip, &tn);
                if (t > 0.0f && t_min > t) { t_min = t; materialid = sc->faces[f].materialid; ip = tip; normal = tn; }
            }
        }
        if (materialid == -1) {
            isx[i].t = -1.0f;
        } else {
            isx[i].t = t_min;
            isx[i].materialId = materialid;
            isx[i].surfaceNormal = normalize(normal);
            isx[i].is_inside = !outside;
            isx[i].intersect = ip;
        }
        if (depth == 0 && iter == 1 && isx[i].t >= 0) {                                  /* :295-304, x-mirrored */
            int col = i % width, row = i / width;
            size_t o = (size_t)(width - col - 1) + (size_t)row * width;
            tensor[(size_t)n * 3 + o] = normal.x;
            tensor[(size_t)n * 4 + o] = normal.y;
            tensor[(size_t)n * 5 + o] = normal.z;
            tensor[(size_t)n * 6 + o] = isx[i].t;
        }
    }
}

static void shade(int iter, int n, const Isect* isx, PathSegment* paths, const SceneView* sc, float* tensor,
                  int depth, int width) {                                                 /* pathtrace.cu:333-390 */
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < n; ++idx) {
        if (paths[idx].remainingBounces == 0) continue;
        Isect in = isx[idx];
        if (in.t > 0.0f) {
            Rng rng = make_rng(iter, idx, paths[idx].remainingBounces);
            Material m = sc->materials[in.materialId];
            if (m.emittance > 0.0f) {
                paths[idx].remainingBounces = 0;
                paths[idx].color = muls(mulv(paths[idx].color, m.color), m.emittance);
            } else {
                scatterRay(&paths[idx], &in, &m, &rng);
                --paths[idx].remainingBounces;
            }
        } else {
            paths[idx].color = V(0, 0, 0);
            paths[idx].remainingBounces = 0;
        }
        if (depth == 0 && iter == 1 && in.t >= 0) {                                       /* :379-387 */
            int col = idx % width, row = idx / width;
            size_t o = (size_t)(width - col - 1) + (size_t)row * width;
            tensor[(size_t)n * 7 + o] = paths[idx].color.x;
            tensor[(size_t)n * 8 + o] = paths[idx].color.y;
            tensor[(size_t)n * 9 + o] = paths[idx].color.z;
        }
    }
}

/* thrust::partition, CUDA back end (pathtrace.cu:505): kept items stable, rejected items reversed */
static int partition_paths(PathSegment* paths, int n, PathSegment* tmp) {
    int k = 0, r = n;
    for (int i = 0; i < n; ++i) {
        if (paths[i].remainingBounces > 0) tmp[k++] = paths[i];
        else tmp[--r] = paths[i];
    }
    memcpy(paths, tmp, sizeof(PathSegment) * (size_t)n);
    return k;
}

/* thrust::sort_by_key with sort_cmp (pathtrace.cu:412-417, :509): stable merge sort on materialId of
 * the UN-compacted intersection slots [0, n) */
static void sort_by_material(PathSegment* paths, Isect* isx, int n, int nmaterials, PathSegment* tmp) {
    int lo = 0, hi = 0;
    for (int i = 0; i < n; ++i) { if (isx[i].materialId < lo) lo = isx[i].materialId; if (isx[i].materialId > hi) hi = isx[i].materialId; }
    (void)nmaterials;
    int nb = hi - lo + 1;
    int* cnt = (int*)calloc((size_t)nb + 1, sizeof(int));
    for (int i = 0; i < n; ++i) cnt[isx[i].materialId - lo + 1]++;
    for (int b = 0; b < nb; ++b) cnt[b + 1] += cnt[b];
    Isect* itmp = (Isect*)malloc(sizeof(Isect) * (size_t)(n ? n : 1));
    for (int i = 0; i < n; ++i) { int d = cnt[isx[i].materialId - lo]++; tmp[d] = paths[i]; itmp[d] = isx[i]; }
    memcpy(paths, tmp, sizeof(PathSegment) * (size_t)n);
    memcpy(isx, itmp, sizeof(Isect) * (size_t)n);
    free(itmp); free(cnt);
}

/* trace callback: called once per bounce after intersect, before shade (the reference's sync point) */
typedef void (*pto_trace_fn)(void* user, int depth, int n, const PathSegment* paths, const Isect* isx);

/* One 1-spp iteration, pathtrace.cu:422-528 plus the per-frame zeroing done by pathtraceInit
 * (:102, :119).  host_tensor: float[10*P].  image: float[3*P] or NULL.  final_paths: PathSegment[P] or NULL.
 * live_counts: int[trace_depth] or NULL.  Returns the sum of live paths over the bounces. */
long long pto_render(int ngeoms, const Geom* geoms, int nmaterials, const Material* materials, int nfaces,
                     const Face* faces, const MeshBox* mesh_box, const Camera* cam, int trace_depth, int iter,
                     int sort_material, float* host_tensor, float* image_out, PathSegment* final_paths,
                     int* live_counts, pto_trace_fn trace, void* user) {
    SceneView sc = {ngeoms, nmaterials, nfaces, trace_depth, geoms, materials, faces, mesh_box};
    const int W = cam->resx, H = cam->resy, P = W * H;
    PathSegment* paths = (PathSegment*)calloc((size_t)P, sizeof(PathSegment));          /* decision D4 */
    PathSegment* tmp = (PathSegment*)malloc(sizeof(PathSegment) * (size_t)P);
    Isect* isx = (Isect*)malloc(sizeof(Isect) * (size_t)P);
    v3* image = (v3*)calloc((size_t)P, sizeof(v3));
    memset(host_tensor, 0, sizeof(float) * 10 * (size_t)P);
    raygen(cam, iter, trace_depth, paths);
    int depth = 0, n = P, done = 0;
    long long sum = 0;
    while (!done) {
        memset(isx, 0, sizeof(Isect) * (size_t)P);                                      /* :478 */
        intersect(depth, n, paths, &sc, host_tensor, isx, iter, W);
        if (trace) trace(user, depth, n, paths, isx);
        if (live_counts) live_counts[depth] = n;
        sum += n;
        shade(iter, n, isx, paths, &sc, host_tensor, depth, W);
        depth++;
        n = partition_paths(paths, n, tmp);                                             /* STREAM_COMPACTION true */
        if (sort_material) sort_by_material(paths, isx, n, nmaterials, tmp);
        done = (n == 0 || depth == trace_depth);
    }
    for (int i = 0; i < P; ++i) {                                                        /* finalGather :393-402 */
        v3* px = &image[paths[i].pixelIndex];
        *px = add(*px, paths[i].color);
    }
    for (int y = 0; y < H; ++y)                                                          /* copy_data :81-94 */
        for (int x = 0; x < W; ++x) {
            v3 pix = image[(W - x - 1) + y * W];
            size_t d = (size_t)x + (size_t)y * W;
            host_tensor[d] = pix.x / (float)iter;
            host_tensor[d + (size_t)P] = pix.y / (float)iter;
            host_tensor[d + 2 * (size_t)P] = pix.z / (float)iter;
        }
    if (image_out) memcpy(image_out, image, sizeof(v3) * (size_t)P);
    if (final_paths) memcpy(final_paths, paths, sizeof(PathSegment) * (size_t)P);
    free(paths); free(tmp); free(isx); free(image);
    return sum;
}

/* micro known-answer hooks */
uint32_t pto_hash_seed(int iter, int index, int depth) { return utilhash((1u << 31) | ((uint32_t)depth << 22) | (uint32_t)iter) ^ utilhash((uint32_t)index); }
void pto_rng_draws(int iter, int index, int depth, int n, float* out) {
    Rng r = make_rng(iter, index, depth);
    for (int i = 0; i < n; ++i) out[i] = rng_uniform(&r, 0, 1);
}
void pto_hemisphere(const float* normal, int iter, int index, int depth, float* out) {
    Rng r = make_rng(iter, index, depth);
    v3 d = calculateRandomDirectionInHemisphere(V(normal[0], normal[1], normal[2]), &r);
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
}
float pto_triangle(const Face* f, const float* o, const float* d, float* ip, float* nrm) {
    Ray r = {V(o[0], o[1], o[2]), V(d[0], d[1], d[2])};
    v3 p = V(0, 0, 0), n = V(0, 0, 0);
    float t = triangleIntersectionTest(f, r, &p, &n);
    ip[0] = p.x; ip[1] = p.y; ip[2] = p.z; nrm[0] = n.x; nrm[1] = n.y; nrm[2] = n.z;
    return t;
}
