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

// Bit-exactness recipe.  The reference's numerics are fixed by TWO contraction steps of its nvcc 12.9 default build:
//   1. NVVM turns some a*b + c into `fma.rn.f32` in the PTX (which product of a sum gets fused depends on basic-block
//      context, so it is not derivable from the source alone), and
//   2. ptxas may still fuse the remaining `mul.f32` / `add.f32` / `sub.f32` (no rounding modifier => contraction allowed)
//      into FFMA in the SASS.
// Step 1 is pinned here by emitting, through inline PTX that NVVM cannot re-associate, exactly the instruction the
// reference's PTX has at each place (`fma.rn.f32` where it fused, plain `mul/add/sub.f32` where it did not); step 2 is then
// left to the same ptxas, which sees the same dataflow and makes the same choice.
__device__ __forceinline__ float fmul(float a, float b) { float r; asm("mul.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm("add.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fsub(float a, float b) { float r; asm("sub.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 add(v3 a, v3 b) { return V(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
__device__ __forceinline__ v3 sub(v3 a, v3 b) { return V(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
__device__ __forceinline__ v3 mulv(v3 a, v3 b) { return V(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
__device__ __forceinline__ v3 muls(v3 a, float s) { return V(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
__device__ __forceinline__ v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
// a + b*s per component, fused (one rounding)
__device__ __forceinline__ v3 fmas(v3 b, float s, v3 a) { return V(ffma(b.x, s, a.x), ffma(b.y, s, a.y), ffma(b.z, s, a.z)); }
// func_geometric.inl:66-73 is tmp = a*b; tmp.x + tmp.y + tmp.z, which the reference's PTX holds as fma(z, z', fma(x, x', y*y'))
__device__ __forceinline__ float dot(v3 a, v3 b) { return ffma(a.z, b.z, ffma(a.x, b.x, fmul(a.y, b.y))); }
// ... except at the first reflect() of scatterRay (interactions.h:219), where it is fma(z, z', fma(y, y', x*x'))
__device__ __forceinline__ float dot_xyz(v3 a, v3 b) { return ffma(a.z, b.z, ffma(a.y, b.y, fmul(a.x, b.x))); }
__device__ __forceinline__ v3 cross(v3 x, v3 y) {                                                       // :134-143, never fused
    return V(fsub(fmul(x.y, y.z), fmul(y.y, x.z)), fsub(fmul(x.z, y.x), fmul(y.z, x.x)), fsub(fmul(x.x, y.y), fmul(y.x, x.y)));
}
__device__ __forceinline__ v3 normalize(v3 a) { return muls(a, __frcp_rn(__fsqrt_rn(dot(a, a)))); }        // :153-159: v * (1 / sqrt(dot))
__device__ __forceinline__ float length(v3 a) { return __fsqrt_rn(dot(a, a)); }                          // :95-101
// reflect(I, N) = I - N * dot(N, I) * 2 (:174-179); d is passed in because its dot pattern differs per call site
__device__ __forceinline__ v3 reflect_d(v3 I, v3 N, float d) { return sub(I, muls(muls(N, d), 2.0f)); }
__device__ __forceinline__ v3 glm_refract(v3 I, v3 N, float eta) {                                       // :192-200
    float d = dot(N, I);
    float k = fsub(1.0f, fmul(fmul(eta, eta), fsub(1.0f, fmul(d, d))));
    v3 r = sub(muls(I, eta), muls(N, ffma(eta, d, __fsqrt_rn(k))));
    return muls(r, (k >= 0.0f) ? 1.0f : 0.0f);
}
// mat4 * vec4 -> xyz (detail/type_mat4x4.inl:617-628): (m0*x + m4*y) + (m8*z + m12*w), column major
__device__ __forceinline__ v3 mulMV1(const float* m, v3 v) {                                             // w = 1
    return V(fadd(ffma(v.x, m[0], fmul(v.y, m[4])), ffma(v.z, m[8], m[12])),
             fadd(ffma(v.x, m[1], fmul(v.y, m[5])), ffma(v.z, m[9], m[13])),
             fadd(ffma(v.x, m[2], fmul(v.y, m[6])), ffma(v.z, m[10], m[14])));
}
__device__ __forceinline__ v3 mulMV0(const float* m, v3 v) {                                             // w = 0 (m12*0 is kept: it can be NaN/-0)
    return V(fadd(ffma(v.x, m[0], fmul(v.y, m[4])), ffma(v.z, m[8], fmul(m[12], 0.0f))),
             fadd(ffma(v.x, m[1], fmul(v.y, m[5])), ffma(v.z, m[9], fmul(m[13], 0.0f))),
             fadd(ffma(v.x, m[2], fmul(v.y, m[6])), ffma(v.z, m[10], fmul(m[14], 0.0f))));
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
    result = fmul(result, 4.656612873077392578125e-10f);      // / 2^31, exact
    return ffma(result, fsub(hi, lo), lo);
}

// ---- intersections.h ------------------------------------------------------------------------------------
__device__ __forceinline__ v3 getPointOnRay(Ray r, float t) {                             // :27-29
    return fmas(normalize(r.d), fadd(t, -.0001f), r.o);
}
__device__ __forceinline__ float comp(const v3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

__device__ inline float boxIntersectionTest(const ptd_geom* box, Ray r, v3& ip, v3& normal, bool& outside) {   // :52-94
    Ray q;
    q.o = mulMV1(box->inverseTransform, r.o);
    q.d = normalize(mulMV0(box->inverseTransform, r.d));
    float tmin = -1e38f, tmax = 1e38f;
    v3 tmin_n = V(0, 0, 0), tmax_n = V(0, 0, 0);
#pragma unroll
    for (int xyz = 0; xyz < 3; ++xyz) {
        float qdxyz = comp(q.d, xyz);
        float t1 = __fdiv_rn(fsub(-0.5f, comp(q.o, xyz)), qdxyz);
        float t2 = __fdiv_rn(fsub(+0.5f, comp(q.o, xyz)), qdxyz);
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
        ip = mulMV1(box->transform, getPointOnRay(q, tmin));
        normal = normalize(mulMV0(box->transform, tmin_n));
        return length(sub(r.o, ip));
    }
    return -1;
}

__device__ inline float sphereIntersectionTest(const ptd_geom* sphere, Ray r, v3& ip, v3& normal, bool& outside) {   // :106-148
    Ray rt;
    rt.o = mulMV1(sphere->inverseTransform, r.o);
    rt.d = normalize(mulMV0(sphere->inverseTransform, r.d));
    float vDotDirection = dot(rt.o, rt.d);
    // v*v - (dot - r^2) sits in the reference's PTX as fma(v, v, r^2 - dot), r^2 = pow(.5, 2)
    float radicand = ffma(vDotDirection, vDotDirection, fsub(0.25f, dot(rt.o, rt.o)));
    if (radicand < 0) return -1;
    float squareRoot = __fsqrt_rn(radicand);
    float firstTerm = -vDotDirection;
    float t1 = fadd(firstTerm, squareRoot);
    float t2 = fsub(firstTerm, squareRoot);
    float t = 0;
    if (t1 < 0 && t2 < 0) return -1;
    else if (t1 > 0 && t2 > 0) { t = fminf(t1, t2); outside = true; }
    else { t = fmaxf(t1, t2); outside = false; }
    v3 obj = getPointOnRay(rt, t);
    ip = mulMV1(sphere->transform, obj);
    normal = normalize(mulMV0(sphere->invTranspose, obj));
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
    float f = __frcp_rn(a);
    v3 s = sub(r.o, v0);
    bx = fmul(f, dot(s, p));
    if (bx < 0.0f) return -1;
    if (bx > 1.0f) return -1;
    v3 q = cross(s, e1);
    by = fmul(f, dot(r.d, q));
    if (by < 0.0f) return -1;
    if (fadd(by, bx) > 1.0f) return -1;
    float bz = fmul(f, dot(e2, q));
    if (!(bz >= 0.0f)) return -1;
    return bz;
}
// intersections.h:159-172: point uses (x, y, 1-x-y) on (v0, v1, v2) (sic), normal uses (1-x-y, x, y) on (n0, n1, n2)
__device__ __forceinline__ void triangleFinish(const ptd_face* f, float bx, float by, v3& ip, v3& normal) {
    const float w = fsub(fsub(1.0f, bx), by);
    ip = fmas(f->v[2], w, fmas(f->v[0], bx, muls(f->v[1], by)));               // (v0*bx + v1*by) + v2*w
    normal = normalize(fmas(f->n[2], by, fmas(f->n[0], w, muls(f->n[1], bx))));  // (n0*w + n1*bx) + n2*by
}

__device__ __forceinline__ bool RayAABBintersect(const Ray& ray, const ptd_aabb& b) {     // :175-200
    float dx = __frcp_rn(ray.d.x), dy = __frcp_rn(ray.d.y), dz = __frcp_rn(ray.d.z);
    float t1 = fmul(fsub(b.lb.x, ray.o.x), dx), t2 = fmul(fsub(b.ub.x, ray.o.x), dx);
    float t3 = fmul(fsub(b.lb.y, ray.o.y), dy), t4 = fmul(fsub(b.ub.y, ray.o.y), dy);
    float t5 = fmul(fsub(b.lb.z, ray.o.z), dz), t6 = fmul(fsub(b.ub.z, ray.o.z), dz);
    float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    if (tmax < 0) return false;
    if (tmin > tmax) return false;
    return true;
}

// ---- interactions.h ---------------------------------------------------------------------------------------
__device__ inline v3 calculateRandomDirectionInHemisphere(v3 normal, Rng& rng) {         // :13-44
    float up = __fsqrt_rn(rng_uniform(rng, 0, 1));
    float over = __fsqrt_rn(fsub(1.0f, fmul(up, up)));
    float around = fmul(rng_uniform(rng, 0, 1), PT_TWO_PI);
    v3 directionNotNormal;
    if (fabsf(normal.x) < PT_SQRT_OF_ONE_THIRD) directionNotNormal = V(1, 0, 0);
    else if (fabsf(normal.y) < PT_SQRT_OF_ONE_THIRD) directionNotNormal = V(0, 1, 0);
    else directionNotNormal = V(0, 0, 1);
    v3 p1 = normalize(cross(normal, directionNotNormal));
    v3 p2 = normalize(cross(normal, p1));
    // (N*up + p1*(cos*over)) + p2*(sin*over): N*up stays a product, the other two terms are fused onto it
    return fmas(p2, fmul(sinf(around), over), fmas(p1, fmul(cosf(around), over), muls(normal, up)));
}
__device__ inline bool refract_hw(v3 v, v3 n, float ni_over_nt, v3& refracted) {         // :74-85
    v3 uv = normalize(v);
    float dt = dot(uv, n);
    // `1.0 - float` is evaluated in double there; rounding that difference to float equals the fp32 subtraction
    float discriminat = fsub(1.0f, fmul(fmul(ni_over_nt, ni_over_nt), fsub(1.0f, fmul(dt, dt))));
    if (discriminat > 0) {
        refracted = sub(muls(sub(uv, muls(n, dt)), ni_over_nt), muls(n, __fsqrt_rn(discriminat)));
        return true;
    }
    return false;
}
__device__ inline float schlick(float cosine, float ref_idx) {                           // :116-120
    float r0 = __fdiv_rn(fsub(1.0f, ref_idx), fadd(ref_idx, 1.0f));
    r0 = fmul(r0, r0);
    // pow(float, int) promotes to double; r0 + (1 - r0) * pow is one fma.rn.f64 in the reference's PTX
    return (float)__fma_rn(pow((double)fsub(1.0f, cosine), 5.0), (double)fsub(1.0f, r0), (double)r0);
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
            refrac_index_ratio = __frcp_rn(m.indexOfRefraction);
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
            dir = normalize(reflect_d(dir, isx_normal, dot_xyz(isx_normal, dir)));
            color = m.specular_color;
        } else {
            dir = normalize(glm_refract(ray.d, normal, refrac_index_ratio));
            if (!length(dir)) {
                dir = normalize(reflect_d(dir, isx_normal, dot(isx_normal, dir)));
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
    ray.o = fmas(dir, 0.01f, isx_point);
    pcolor = mulv(pcolor, color);
}
}  // namespace ptm
