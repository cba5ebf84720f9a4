// Device math of the path-trace hot path (HP-1).
//
// Restates, as scalar fp32 expression trees, what the reference's kernels evaluate through GLM 0.9.6.3 and
// thrust::minstd_rand: Inference/src/intersections.h, interactions.h, pathtrace.cu:52-56.  The association
// order of every sum/product mirrors GLM's (dot = tmp.x+tmp.y+tmp.z, normalize = v * (1/sqrt(dot)), mat4*vec4 =
// (m0*v0+m1*v1)+(m2*v2+m3*v3), ...) and the file is compiled with nvcc's defaults (-fmad=true, precise
// div/sqrt, libdevice sinf/cosf/pow) exactly like the reference's own build (SURVEY.md decision D6), so that
// hit/miss decisions - and with them the compacted PathSegment order - come out bit-identical.
#pragma once
#include <cfloat>
#include <cstdint>
#include "ptd.h"

#define PT_TWO_PI 6.2831853071795864769252867665590057683943f          // utilities.h:14
#define PT_SQRT_OF_ONE_THIRD 0.5773502691896257645091487805019574556476f   // utilities.h:15

namespace ptm {
typedef ptd_vec3 v3;
struct Ray { v3 o, d; };

__device__ __forceinline__ v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(v3 a, v3 b) { v3 t = mulv(a, b); return t.x + t.y + t.z; }            // func_geometric.inl:66-73
__device__ __forceinline__ v3 cross(v3 x, v3 y) {                                                       // :134-143
    return V(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
__device__ __forceinline__ v3 normalize(v3 a) { return muls(a, 1.0f / sqrtf(dot(a, a))); }                // :153-159
__device__ __forceinline__ float length(v3 a) { return sqrtf(dot(a, a)); }                               // :95-101
__device__ __forceinline__ v3 reflect(v3 I, v3 N) { return sub(I, muls(muls(N, dot(N, I)), 2.0f)); }       // :174-179
__device__ __forceinline__ v3 glm_refract(v3 I, v3 N, float eta) {                                       // :192-200
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    v3 r = sub(muls(I, eta), muls(N, eta * d + sqrtf(k)));
    return muls(r, (float)(k >= 0.0f));
}
// mat4 * vec4 -> xyz (detail/type_mat4x4.inl:617-628), column major
__device__ __forceinline__ v3 mulMV(const float* m, v3 v, float w) {
    v3 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * w);
    return r;
}
__device__ __forceinline__ float glm_min(float x, float y) { return x < y ? x : y; }      // as pinned by oracle/pt_oracle.c
__device__ __forceinline__ float glm_max(float x, float y) { return x > y ? x : y; }

// ---- RNG: intersections.h:12-20, pathtrace.cu:52-56, thrust::minstd_rand + uniform_real_distribution ----
__device__ __forceinline__ uint32_t utilhash(uint32_t a) {
    a = (a + 0x7ed55d16) + (a << 12);
    a = (a ^ 0xc761c23c) ^ (a >> 19);
    a = (a + 0x165667b1) + (a << 5);
    a = (a + 0xd3a2646c) ^ (a << 9);
    a = (a + 0xfd7046c5) + (a << 3);
    a = (a ^ 0xb55a4f09) ^ (a >> 16);
    return a;
}
struct Rng { uint32_t x; };
__device__ __forceinline__ Rng make_rng(int iter, int index, int depth) {
    uint32_t h = utilhash((1u << 31) | ((uint32_t)depth << 22) | (uint32_t)iter) ^ utilhash((uint32_t)index);
    Rng r;
    r.x = h % 2147483647u;                       // linear_congruential_engine.inl:45-56
    if (r.x == 0) r.x = 1;
    return r;
}
__device__ __forceinline__ uint32_t rng_next(Rng& r) {   // x <- 48271 x mod (2^31-1), Schrage form (thrust/random/detail/mod.h)
    const uint32_t a = 48271u, m = 2147483647u, q = m / a, rr = m % a;
    uint32_t x = r.x;
    uint32_t t1 = a * (x % q), t2 = rr * (x / q);
    x = (t1 >= t2) ? (t1 - t2) : (m - t2 + t1);
    r.x = x;
    return x;
}
__device__ __forceinline__ float rng_uniform(Rng& r, float lo, float hi) {   // uniform_real_distribution.inl:60-73
    float result = (float)(rng_next(r) - 1u);
    result /= (1.0f + (float)(2147483646u - 1u));
    return (result * (hi - lo)) + lo;
}

// ---- intersections.h ------------------------------------------------------------------------------------
__device__ __forceinline__ v3 getPointOnRay(Ray r, float t) {                             // :27-29
    return add(r.o, muls(normalize(r.d), (t - .0001f)));
}
__device__ __forceinline__ float comp(const v3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

__device__ inline float boxIntersectionTest(const ptd_geom* box, Ray r, v3& ip, v3& normal, bool& outside) {   // :52-94
    Ray q;
    q.o = mulMV(box->inverseTransform, r.o, 1.0f);
    q.d = normalize(mulMV(box->inverseTransform, r.d, 0.0f));
    float tmin = -1e38f, tmax = 1e38f;
    v3 tmin_n = V(0, 0, 0), tmax_n = V(0, 0, 0);
#pragma unroll
    for (int xyz = 0; xyz < 3; ++xyz) {
        float qdxyz = comp(q.d, xyz);
        float t1 = (-0.5f - comp(q.o, xyz)) / qdxyz;
        float t2 = (+0.5f - comp(q.o, xyz)) / qdxyz;
        float ta = glm_min(t1, t2);
        float tb = glm_max(t1, t2);
        float s = t2 < t1 ? +1.0f : -1.0f;
        v3 n = V(xyz == 0 ? s : 0.f, xyz == 1 ? s : 0.f, xyz == 2 ? s : 0.f);
        if (ta > 0 && ta > tmin) { tmin = ta; tmin_n = n; }
        if (tb < tmax) { tmax = tb; tmax_n = n; }
    }
    if (tmax >= tmin && tmax > 0) {
        outside = true;
        if (tmin <= 0) { tmin = tmax; tmin_n = tmax_n; outside = false; }
        ip = mulMV(box->transform, getPointOnRay(q, tmin), 1.0f);
        normal = normalize(mulMV(box->transform, tmin_n, 0.0f));
        return length(sub(r.o, ip));
    }
    return -1;
}

__device__ inline float sphereIntersectionTest(const ptd_geom* sphere, Ray r, v3& ip, v3& normal, bool& outside) {   // :106-148
    Ray rt;
    rt.o = mulMV(sphere->inverseTransform, r.o, 1.0f);
    rt.d = normalize(mulMV(sphere->inverseTransform, r.d, 0.0f));
    float vDotDirection = dot(rt.o, rt.d);
    float radicand = vDotDirection * vDotDirection - (dot(rt.o, rt.o) - 0.25f /* pow(radius = .5, 2) */);
    if (radicand < 0) return -1;
    float squareRoot = sqrtf(radicand);
    float firstTerm = -vDotDirection;
    float t1 = firstTerm + squareRoot;
    float t2 = firstTerm - squareRoot;
    float t = 0;
    if (t1 < 0 && t2 < 0) return -1;
    else if (t1 > 0 && t2 > 0) { t = fminf(t1, t2); outside = true; }
    else { t = fmaxf(t1, t2); outside = false; }
    v3 obj = getPointOnRay(rt, t);
    ip = mulMV(sphere->transform, obj, 1.0f);
    normal = normalize(mulMV(sphere->invTranspose, obj, 0.0f));
    if (!outside) normal = neg(normal);
    return length(sub(r.o, ip));
}

// glm/gtx/intersect.inl:37-74 (back-face culling Moeller-Trumbore).  Returns the ray parameter t (bary.z) or -1;
// bx, by are the first two "barycentrics" the reference then feeds to its (mis-mapped) point interpolation.
__device__ __forceinline__ float triangleParam(v3 v0, v3 v1, v3 v2, Ray r, float& bx, float& by) {
    v3 e1 = sub(v1, v0);
    v3 e2 = sub(v2, v0);
    v3 p = cross(r.d, e2);
    float a = dot(e1, p);
    if (a < FLT_EPSILON) return -1;
    float f = 1.0f / a;
    v3 s = sub(r.o, v0);
    bx = f * dot(s, p);
    if (bx < 0.0f) return -1;
    if (bx > 1.0f) return -1;
    v3 q = cross(s, e1);
    by = f * dot(r.d, q);
    if (by < 0.0f) return -1;
    if (by + bx > 1.0f) return -1;
    float bz = f * dot(e2, q);
    if (!(bz >= 0.0f)) return -1;
    return bz;
}
// intersections.h:159-172: point uses (x, y, 1-x-y) on (v0, v1, v2) (sic), normal uses (1-x-y, x, y) on (n0, n1, n2)
__device__ __forceinline__ void triangleFinish(const ptd_face* f, float bx, float by, v3& ip, v3& normal) {
    ip = add(add(muls(f->v[0], bx), muls(f->v[1], by)), muls(f->v[2], (1 - bx - by)));
    normal = normalize(add(add(muls(f->n[0], (1 - bx - by)), muls(f->n[1], bx)), muls(f->n[2], by)));
}

__device__ __forceinline__ bool RayAABBintersect(const Ray& ray, const ptd_aabb& b) {     // :175-200
    float dx = 1.0f / ray.d.x, dy = 1.0f / ray.d.y, dz = 1.0f / ray.d.z;
    float t1 = (b.lb.x - ray.o.x) * dx, t2 = (b.ub.x - ray.o.x) * dx;
    float t3 = (b.lb.y - ray.o.y) * dy, t4 = (b.ub.y - ray.o.y) * dy;
    float t5 = (b.lb.z - ray.o.z) * dz, t6 = (b.ub.z - ray.o.z) * dz;
    float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    if (tmax < 0) return false;
    if (tmin > tmax) return false;
    return true;
}

// ---- interactions.h ---------------------------------------------------------------------------------------
__device__ inline v3 calculateRandomDirectionInHemisphere(v3 normal, Rng& rng) {         // :13-44
    float up = sqrtf(rng_uniform(rng, 0, 1));
    float over = sqrtf(1 - up * up);
    float around = rng_uniform(rng, 0, 1) * PT_TWO_PI;
    v3 directionNotNormal;
    if (fabsf(normal.x) < PT_SQRT_OF_ONE_THIRD) directionNotNormal = V(1, 0, 0);
    else if (fabsf(normal.y) < PT_SQRT_OF_ONE_THIRD) directionNotNormal = V(0, 1, 0);
    else directionNotNormal = V(0, 0, 1);
    v3 p1 = normalize(cross(normal, directionNotNormal));
    v3 p2 = normalize(cross(normal, p1));
    return add(add(muls(normal, up), muls(p1, cosf(around) * over)), muls(p2, sinf(around) * over));
}
__device__ inline bool refract_hw(v3 v, v3 n, float ni_over_nt, v3& refracted) {         // :74-85
    v3 uv = normalize(v);
    float dt = dot(uv, n);
    float discriminat = (float)(1.0 - (double)(ni_over_nt * ni_over_nt * (1 - dt * dt)));   // `1.0` is a double literal there
    if (discriminat > 0) {
        refracted = sub(muls(sub(uv, muls(n, dt)), ni_over_nt), muls(n, sqrtf(discriminat)));
        return true;
    }
    return false;
}
__device__ inline float schlick(float cosine, float ref_idx) {                           // :116-120
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    return (float)((double)r0 + (double)(1 - r0) * pow((double)(1 - cosine), 5.0));       // pow(float,int) promotes to double
}
// live branch of scatterRay under the reference's default macros: interactions.h:194-258
__device__ inline void scatterRay(Ray& ray, v3& pcolor, v3 isx_point, v3 isx_normal, const ptd_material& m, Rng& rng) {
    v3 dir = ray.d;
    v3 color = V(1.0f, 1.0f, 1.0f);
    float reflective_prob = m.hasReflective;
    if (reflective_prob != 0 || m.hasRefractive != 0) {
        float pdf = rng_uniform(rng, 0, 1), refrac_index_ratio, cosine;
        v3 normal;
        cosine = dot(normalize(dir), isx_normal);
        if (cosine <= 0) {
            normal = isx_normal;
            refrac_index_ratio = 1 / m.indexOfRefraction;
            cosine = -cosine;
        } else {
            normal = neg(isx_normal);
            refrac_index_ratio = m.indexOfRefraction;
        }
        if (refract_hw(ray.d, normal, refrac_index_ratio, dir))
            reflective_prob = schlick(cosine, refrac_index_ratio);
        else
            reflective_prob = 1.0f;
        if (pdf < reflective_prob) {
            dir = normalize(reflect(dir, isx_normal));
            color = m.specular_color;
        } else {
            dir = normalize(glm_refract(ray.d, normal, refrac_index_ratio));
            if (!length(dir)) {
                dir = normalize(reflect(dir, isx_normal));
                color = m.specular_color;
            } else {
                color = m.color;
            }
        }
    } else {
        dir = normalize(calculateRandomDirectionInHemisphere(isx_normal, rng));
        color = m.color;
    }
    ray.d = dir;
    ray.o = add(isx_point, muls(dir, 0.01f));
    pcolor = mulv(pcolor, color);
}
}  // namespace ptm
