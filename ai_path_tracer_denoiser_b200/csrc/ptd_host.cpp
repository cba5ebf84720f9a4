// Host side of libptd.so: error channel, scene-file + OBJ ingest, camera maths, BVH build.
//
// Scene grammar and quirks follow Inference/src/scene.cpp:11-320 and utilities.cpp:45-92 of the reference
// (SURVEY.md section 8f-1); the matrix algebra restates GLM 0.9.6.3's expression order
// (gtc/matrix_transform.inl:40-134, detail/type_mat4x4.inl:37-92,686-704, gtc/matrix_inverse.inl:94-147) so
// that the Geom records are bit-identical to what the reference's loader produces.  Nothing here is on the
// per-frame path.  Compiled with -ffp-contract=off (x86-64 host code of the reference has no FMA either).
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <algorithm>
#include "ptd_internal.h"

// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void ptd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* ptd_last_error(void) { return g_err; }
extern "C" int ptd_version(void) { return 100; }
extern "C" int ptd_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(ptd_path_segment);
        case 1: return (int)sizeof(ptd_intersection);
        case 2: return (int)sizeof(ptd_geom);
        case 3: return (int)sizeof(ptd_face);
        case 4: return (int)sizeof(ptd_material);
        case 5: return (int)sizeof(ptd_camera);
        case 6: return (int)sizeof(ptd_aabb);
    }
    return -1;
}

// ---- GLM-shaped fp32 algebra (column-major mat4: m[c*4 + r]) ----------------------------------------
#define PI_F 3.1415926535897932384626422832795028841971f   // utilities.h:13
namespace {
struct V4 { float x, y, z, w; };
inline V4 mul(V4 a, float s) { return V4{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 add(V4 a, V4 b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 sub(V4 a, V4 b) { return V4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline V4 mulv(V4 a, V4 b) { return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
struct M4 {
    float m[16];
    V4 col(int c) const { return V4{m[c * 4], m[c * 4 + 1], m[c * 4 + 2], m[c * 4 + 3]}; }
    void set(int c, V4 v) { m[c * 4] = v.x; m[c * 4 + 1] = v.y; m[c * 4 + 2] = v.z; m[c * 4 + 3] = v.w; }
    float at(int c, int r) const { return m[c * 4 + r]; }
};
M4 identity() { M4 r; memset(r.m, 0, sizeof r.m); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }

inline ptd_vec3 v3(float x, float y, float z) { ptd_vec3 r = {x, y, z}; return r; }
inline float dot3(ptd_vec3 a, ptd_vec3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
inline ptd_vec3 norm3(ptd_vec3 a) { float s = 1.0f / sqrtf(dot3(a, a)); return v3(a.x * s, a.y * s, a.z * s); }
inline ptd_vec3 cross3(ptd_vec3 x, ptd_vec3 y) {
    return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
inline ptd_vec3 sub3(ptd_vec3 a, ptd_vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline ptd_vec3 add3(ptd_vec3 a, ptd_vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float len3(ptd_vec3 a) { return sqrtf(dot3(a, a)); }
inline ptd_vec3 muls3(ptd_vec3 a, float k) { return v3(a.x * k, a.y * k, a.z * k); }

M4 matmul(const M4& a, const M4& b) {                       // type_mat4x4.inl:686-704
    M4 r;
    for (int j = 0; j < 4; ++j)
        r.set(j, add(add(add(mul(a.col(0), b.at(j, 0)), mul(a.col(1), b.at(j, 1))), mul(a.col(2), b.at(j, 2))), mul(a.col(3), b.at(j, 3))));
    return r;
}
M4 translate(const M4& m, ptd_vec3 v) {                     // matrix_transform.inl:40-49
    M4 r = m;
    r.set(3, add(add(add(mul(m.col(0), v.x), mul(m.col(1), v.y)), mul(m.col(2), v.z)), m.col(3)));
    return r;
}
M4 rotate(const M4& m, float angle, ptd_vec3 v) {           // matrix_transform.inl:52-85
    const float a = angle, c = cosf(a), s = sinf(a);
    ptd_vec3 axis = norm3(v);
    ptd_vec3 temp = v3((1.f - c) * axis.x, (1.f - c) * axis.y, (1.f - c) * axis.z);
    float R[3][3];
    R[0][0] = c + temp.x * axis.x;
    R[0][1] = 0 + temp.x * axis.y + s * axis.z;
    R[0][2] = 0 + temp.x * axis.z - s * axis.y;
    R[1][0] = 0 + temp.y * axis.x - s * axis.z;
    R[1][1] = c + temp.y * axis.y;
    R[1][2] = 0 + temp.y * axis.z + s * axis.x;
    R[2][0] = 0 + temp.z * axis.x + s * axis.y;
    R[2][1] = 0 + temp.z * axis.y - s * axis.x;
    R[2][2] = c + temp.z * axis.z;
    M4 r;
    for (int j = 0; j < 3; ++j)
        r.set(j, add(add(mul(m.col(0), R[j][0]), mul(m.col(1), R[j][1])), mul(m.col(2), R[j][2])));
    r.set(3, m.col(3));
    return r;
}
M4 scale(const M4& m, ptd_vec3 v) {                         // matrix_transform.inl:122-134
    M4 r;
    r.set(0, mul(m.col(0), v.x)); r.set(1, mul(m.col(1), v.y)); r.set(2, mul(m.col(2), v.z)); r.set(3, m.col(3));
    return r;
}
M4 inverse(const M4& M) {                                   // type_mat4x4.inl:37-92
#define m(c, r) M.at(c, r)
    float Coef00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), Coef02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), Coef03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
    float Coef04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3), Coef06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3), Coef07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
    float Coef08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2), Coef10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2), Coef11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
    float Coef12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3), Coef14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3), Coef15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
    float Coef16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), Coef18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), Coef19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
    float Coef20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1), Coef22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), Coef23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
    V4 Fac0{Coef00, Coef00, Coef02, Coef03}, Fac1{Coef04, Coef04, Coef06, Coef07}, Fac2{Coef08, Coef08, Coef10, Coef11};
    V4 Fac3{Coef12, Coef12, Coef14, Coef15}, Fac4{Coef16, Coef16, Coef18, Coef19}, Fac5{Coef20, Coef20, Coef22, Coef23};
    V4 Vec0{m(1, 0), m(0, 0), m(0, 0), m(0, 0)}, Vec1{m(1, 1), m(0, 1), m(0, 1), m(0, 1)};
    V4 Vec2{m(1, 2), m(0, 2), m(0, 2), m(0, 2)}, Vec3{m(1, 3), m(0, 3), m(0, 3), m(0, 3)};
#undef m
    V4 Inv0 = add(sub(mulv(Vec1, Fac0), mulv(Vec2, Fac1)), mulv(Vec3, Fac2));
    V4 Inv1 = add(sub(mulv(Vec0, Fac0), mulv(Vec2, Fac3)), mulv(Vec3, Fac4));
    V4 Inv2 = add(sub(mulv(Vec0, Fac1), mulv(Vec1, Fac3)), mulv(Vec3, Fac5));
    V4 Inv3 = add(sub(mulv(Vec0, Fac2), mulv(Vec1, Fac4)), mulv(Vec2, Fac5));
    V4 SignA{+1, -1, +1, -1}, SignB{-1, +1, -1, +1};
    M4 Inv;
    Inv.set(0, mulv(Inv0, SignA)); Inv.set(1, mulv(Inv1, SignB)); Inv.set(2, mulv(Inv2, SignA)); Inv.set(3, mulv(Inv3, SignB));
    V4 Row0{Inv.at(0, 0), Inv.at(1, 0), Inv.at(2, 0), Inv.at(3, 0)};
    V4 Dot0 = mulv(M.col(0), Row0);
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    float OneOverDeterminant = 1.0f / Dot1;
    M4 r;
    for (int c = 0; c < 4; ++c) r.set(c, mul(Inv.col(c), OneOverDeterminant));
    return r;
}
M4 inverseTranspose(const M4& M) {                          // gtc/matrix_inverse.inl:94-147
#define m(c, r) M.at(c, r)
    float S00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), S01 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3), S02 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2);
    float S03 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3), S04 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), S05 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1);
    float S06 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), S07 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3), S08 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2);
    float S09 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3), S10 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), S11 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3);
    float S12 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), S13 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3), S14 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
    float S15 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2), S16 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3), S17 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
    float S18 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
    M4 I;
    float* o = I.m;
    o[0] = +(m(1, 1) * S00 - m(1, 2) * S01 + m(1, 3) * S02);
    o[1] = -(m(1, 0) * S00 - m(1, 2) * S03 + m(1, 3) * S04);
    o[2] = +(m(1, 0) * S01 - m(1, 1) * S03 + m(1, 3) * S05);
    o[3] = -(m(1, 0) * S02 - m(1, 1) * S04 + m(1, 2) * S05);
    o[4] = -(m(0, 1) * S00 - m(0, 2) * S01 + m(0, 3) * S02);
    o[5] = +(m(0, 0) * S00 - m(0, 2) * S03 + m(0, 3) * S04);
    o[6] = -(m(0, 0) * S01 - m(0, 1) * S03 + m(0, 3) * S05);
    o[7] = +(m(0, 0) * S02 - m(0, 1) * S04 + m(0, 2) * S05);
    o[8] = +(m(0, 1) * S06 - m(0, 2) * S07 + m(0, 3) * S08);
    o[9] = -(m(0, 0) * S06 - m(0, 2) * S09 + m(0, 3) * S10);
    o[10] = +(m(0, 0) * S11 - m(0, 1) * S09 + m(0, 3) * S12);
    o[11] = -(m(0, 0) * S08 - m(0, 1) * S10 + m(0, 2) * S12);
    o[12] = -(m(0, 1) * S13 - m(0, 2) * S14 + m(0, 3) * S15);
    o[13] = +(m(0, 0) * S13 - m(0, 2) * S16 + m(0, 3) * S17);
    o[14] = -(m(0, 0) * S14 - m(0, 1) * S16 + m(0, 3) * S18);
    o[15] = +(m(0, 0) * S15 - m(0, 1) * S17 + m(0, 2) * S18);
    float Determinant = +m(0, 0) * o[0] + m(0, 1) * o[1] + m(0, 2) * o[2] + m(0, 3) * o[3];
#undef m
    for (int i = 0; i < 16; ++i) o[i] /= Determinant;
    return I;
}
M4 buildTransformationMatrix(ptd_vec3 t, ptd_vec3 r, ptd_vec3 s) {    // utilities.cpp:45-52
    M4 T = translate(identity(), t);
    M4 R = rotate(identity(), r.x * PI_F / 180, v3(1, 0, 0));
    R = matmul(R, rotate(identity(), r.y * PI_F / 180, v3(0, 1, 0)));
    R = matmul(R, rotate(identity(), r.z * PI_F / 180, v3(0, 0, 1)));
    M4 S = scale(identity(), s);
    return matmul(matmul(T, R), S);
}

// ---- text helpers (utilities.cpp:54-92) --------------------------------------------------------------
bool safeGetline(std::istream& is, std::string& t) {
    t.clear();
    std::streambuf* sb = is.rdbuf();
    if (!is.good()) return false;
    for (;;) {
        int c = sb->sbumpc();
        switch (c) {
            case '\n': return true;
            case '\r': if (sb->sgetc() == '\n') sb->sbumpc(); return true;
            case EOF: if (t.empty()) is.setstate(std::ios::eofbit); return true;
            default: t += (char)c;
        }
    }
}
std::vector<std::string> tokenize(const std::string& s) {
    std::istringstream ss(s);
    std::vector<std::string> out;
    std::string w;
    while (ss >> w) out.push_back(w);
    return out;
}
inline float tokf(const std::vector<std::string>& t, size_t i) { return i < t.size() ? (float)atof(t[i].c_str()) : 0.f; }
inline int toki(const std::vector<std::string>& t, size_t i) { return i < t.size() ? atoi(t[i].c_str()) : 0; }
inline ptd_vec3 tok3(const std::vector<std::string>& t) { return v3(tokf(t, 1), tokf(t, 2), tokf(t, 3)); }

// ---- OBJ ingest: the subset of tinyobjloader the reference relies on (scene.cpp:259-317) --------------
// Real numbers follow tinyobj's own digit-accumulating parser (tiny_obj_loader.h:805-929), not strtod, so
// the float bits match: integer digits mantissa*10+d, fraction digits d*10^-k (table for k<8, pow beyond),
// exponent through ldexp(mantissa*5^e, e).
bool parse_real(const char*& p, float* out) {
    while (*p == ' ' || *p == '\t') ++p;
    const char* s = p;
    const char* e = s;
    while (*e && *e != ' ' && *e != '\t' && *e != '\r' && *e != '\n') ++e;
    p = e;
    if (s >= e) return false;
    double mantissa = 0.0;
    int exponent = 0, read = 0;
    char sign = '+', exp_sign = '+';
    const char* c = s;
    bool lead_dot = false;
    auto digit = [](char ch) { return ch >= '0' && ch <= '9'; };
    if (*c == '+' || *c == '-') { sign = *c; ++c; if (c != e && *c == '.') lead_dot = true; }
    else if (digit(*c)) {}
    else if (*c == '.') lead_dot = true;
    else return false;
    if (!lead_dot) {
        while (c != e && digit(*c)) { mantissa *= 10; mantissa += (int)(*c - '0'); ++c; ++read; }
        if (read == 0) return false;
    }
    if (c != e && *c == '.') {
        ++c; read = 1;
        static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
        while (c != e && digit(*c)) {
            mantissa += (int)(*c - '0') * (read < 8 ? lut[read] : std::pow(10.0, -read));
            ++read; ++c;
        }
    }
    if (c != e && (*c == 'e' || *c == 'E')) {
        ++c;
        if (c != e && (*c == '+' || *c == '-')) { exp_sign = *c; ++c; }
        else if (c == e || !digit(*c)) return false;
        read = 0;
        while (c != e && digit(*c)) { exponent = exponent * 10 + (int)(*c - '0'); ++c; ++read; }
        exponent *= (exp_sign == '+' ? 1 : -1);
        if (read == 0) return false;
    }
    double r = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
    *out = (float)r;
    return true;
}
struct ObjIdx { int v, vt, vn; };
bool fix_index(int idx, int n, int* out) {               // tiny_obj_loader.h fixIndex: 1-based, negative = relative
    if (idx > 0) { *out = idx - 1; return true; }
    if (idx == 0) return false;
    *out = n + idx;
    return true;
}
bool parse_triple(const char*& p, int nv, int nt, int nn, ObjIdx* out) {
    while (*p == ' ' || *p == '\t') ++p;
    if (!*p || *p == '\r' || *p == '\n') return false;
    out->v = out->vt = out->vn = -1;
    if (!fix_index(atoi(p), nv, &out->v)) return false;
    p += strcspn(p, "/ \t\r\n");
    if (*p != '/') return true;
    ++p;
    if (*p == '/') {                                     // v//vn
        ++p;
        if (!fix_index(atoi(p), nn, &out->vn)) return false;
        p += strcspn(p, "/ \t\r\n");
        return true;
    }
    if (!fix_index(atoi(p), nt, &out->vt)) return false;  // v/vt[/vn]
    p += strcspn(p, "/ \t\r\n");
    if (*p != '/') return true;
    ++p;
    if (!fix_index(atoi(p), nn, &out->vn)) return false;
    p += strcspn(p, "/ \t\r\n");
    return true;
}

std::string dirname_of(const std::string& p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

// Extension (README "experimental MTL parsing"; the reference loads the OBJ's materials and then ignores them,
// scene.cpp:259-266,314): with `USEMTL 1` in the MESH block, `mtllib` / `usemtl` statements give every face its own material.
struct ObjMtl { std::string name; ptd_material m; };
bool load_mtl(const std::string& path, std::vector<ObjMtl>& out) {
    std::ifstream f(path.c_str());
    if (!f.is_open()) return false;
    std::string line;
    int illum = 2; float ke[3] = {0, 0, 0};
    auto finish = [&]() {
        if (out.empty()) return;
        ptd_material& m = out.back().m;
        // illum 3 / 5: mirror reflection; 4 / 6 / 7: glass (refraction + Fresnel reflection); Ke: emitter
        m.hasReflective = (illum == 3 || illum == 5) ? 1.f : 0.f;
        m.hasRefractive = (illum == 4 || illum == 6 || illum == 7) ? 1.f : 0.f;
        const float e = std::max(ke[0], std::max(ke[1], ke[2]));
        if (e > 0.f) { m.emittance = e; m.color = v3(ke[0] / e, ke[1] / e, ke[2] / e); }
    };
    while (safeGetline(f, line) && (f.good() || !line.empty())) {
        std::vector<std::string> t = tokenize(line);
        if (t.empty()) continue;
        if (t[0] == "newmtl") {
            finish();
            ObjMtl m;
            m.name = t.size() > 1 ? t[1] : "";
            memset(&m.m, 0, sizeof m.m);
            m.m.color = v3(1, 1, 1); m.m.indexOfRefraction = 1.f;
            out.push_back(m);
            illum = 2; ke[0] = ke[1] = ke[2] = 0.f;
        } else if (out.empty()) {
            continue;
        } else if (t[0] == "Kd") out.back().m.color = tok3(t);
        else if (t[0] == "Ks") out.back().m.specular_color = tok3(t);
        else if (t[0] == "Ns") out.back().m.specular_exponent = tokf(t, 1);
        else if (t[0] == "Ni") out.back().m.indexOfRefraction = tokf(t, 1);
        else if (t[0] == "illum") illum = toki(t, 1);
        else if (t[0] == "Ke") { ke[0] = tokf(t, 1); ke[1] = tokf(t, 2); ke[2] = tokf(t, 3); }
    }
    finish();
    return true;
}

// face_mtl (optional): receives, per appended face, the index into `mtls` selected by the last `usemtl`, or -1
ptd_status load_obj(ptd_scene& sc, const std::string& path, int materialid, const M4& transform,
                    std::vector<ObjMtl>* mtls = nullptr, std::vector<int>* face_mtl = nullptr) {
    std::ifstream f(path.c_str());
    if (!f.is_open()) PTD_FAIL(PTD_ERR_IO, "cannot open OBJ file '%s'", path.c_str());
    std::vector<float> V, N;
    std::string line;
    std::vector<ObjIdx> poly;
    long lineno = 0;
    int cur_mtl = -1;
    while (safeGetline(f, line) && (f.good() || !line.empty())) {
        ++lineno;
        const char* p = line.c_str();
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            p += 2;
            float x = 0, y = 0, z = 0;
            parse_real(p, &x); parse_real(p, &y); parse_real(p, &z);
            V.push_back(x); V.push_back(y); V.push_back(z);
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
            p += 3;
            float x = 0, y = 0, z = 0;
            parse_real(p, &x); parse_real(p, &y); parse_real(p, &z);
            N.push_back(x); N.push_back(y); N.push_back(z);
        } else if (mtls && strncmp(p, "mtllib", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
            std::vector<std::string> t = tokenize(line);
            if (t.size() > 1) load_mtl(dirname_of(path) + "/" + t[1], *mtls);
        } else if (mtls && strncmp(p, "usemtl", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
            std::vector<std::string> t = tokenize(line);
            cur_mtl = -1;
            for (size_t k = 0; t.size() > 1 && k < mtls->size(); ++k) if ((*mtls)[k].name == t[1]) cur_mtl = (int)k;
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            p += 2;
            poly.clear();
            ObjIdx ix;
            while (parse_triple(p, (int)V.size() / 3, 0, (int)N.size() / 3, &ix)) poly.push_back(ix);
            if (poly.size() < 3) continue;
            for (size_t k = 2; k < poly.size(); ++k) {     // triangulate=true: fan (i0, i[k-1], i[k])
                const ObjIdx tri[3] = {poly[0], poly[k - 1], poly[k]};
                ptd_face face;
                memset(&face, 0, sizeof face);
                bool have_n = true;
                for (int v = 0; v < 3; ++v) {
                    if (tri[v].v < 0 || tri[v].v * 3 + 2 >= (int)V.size())
                        PTD_FAIL(PTD_ERR_PARSE, "%s:%ld: vertex index out of range", path.c_str(), lineno);
                    V4 pos{V[3 * tri[v].v], V[3 * tri[v].v + 1], V[3 * tri[v].v + 2], 1.f};
                    // transform * vec4 (type_mat4x4.inl:617-628): (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
                    V4 t = add(add(mul(transform.col(0), pos.x), mul(transform.col(1), pos.y)),
                               add(mul(transform.col(2), pos.z), mul(transform.col(3), pos.w)));
                    face.v[v] = v3(t.x, t.y, t.z);
                    if (sc.mesh_box.lb.x > t.x) sc.mesh_box.lb.x = t.x;      // scene.h:28-42
                    if (sc.mesh_box.lb.y > t.y) sc.mesh_box.lb.y = t.y;
                    if (sc.mesh_box.lb.z > t.z) sc.mesh_box.lb.z = t.z;
                    if (sc.mesh_box.ub.x < t.x) sc.mesh_box.ub.x = t.x;
                    if (sc.mesh_box.ub.y < t.y) sc.mesh_box.ub.y = t.y;
                    if (sc.mesh_box.ub.z < t.z) sc.mesh_box.ub.z = t.z;
                    if (tri[v].vn >= 0 && tri[v].vn * 3 + 2 < (int)N.size())
                        face.n[v] = norm3(v3(N[3 * tri[v].vn], N[3 * tri[v].vn + 1], N[3 * tri[v].vn + 2]));   // scene.cpp:304-307 (untransformed)
                    else
                        have_n = false;
                }
                if (!have_n) {   // the reference reads out of bounds here; we fall back to its RECOMPUTE_NORMALS formula (scene.cpp:198-204)
                    ptd_vec3 g = norm3(cross3(sub3(face.v[2], face.v[0]), sub3(face.v[1], face.v[0])));
                    face.n[0] = face.n[1] = face.n[2] = g;
                }
                face.materialid = materialid;
                sc.faces.push_back(face);
                if (face_mtl) face_mtl->push_back(cur_mtl);
            }
        }
    }
    return PTD_OK;
}

}  // namespace

// scene.cpp:142-152: derived camera fields (fov uses tan of the FULL fovy, sic)
void ptd_camera_derive(ptd_camera& cam, float fovy) {
    float yscaled = tanf(fovy * (PI_F / 180));
    float xscaled = (yscaled * cam.res_x) / cam.res_y;
    float fovx = (atanf(xscaled) * 180) / PI_F;
    cam.fov_x = fovx;
    cam.fov_y = fovy;
    cam.pixelLength_x = 2 * xscaled / (float)cam.res_x;
    cam.pixelLength_y = 2 * yscaled / (float)cam.res_y;
}

extern "C" ptd_status ptd_scene_load(const char* path, ptd_scene** out) {
    if (!path || !out) PTD_FAIL(PTD_ERR_ARG, "ptd_scene_load: null argument");
    *out = nullptr;
    std::ifstream in(path);
    if (!in.is_open()) PTD_FAIL(PTD_ERR_IO, "cannot open scene file '%s'", path);   // scene.cpp:16-19 throws here
    ptd_scene* sc = new ptd_scene();
    memset(&sc->mesh_box, 0, sizeof sc->mesh_box);
    memset(&sc->camera, 0, sizeof sc->camera);
    bool have_camera = false;
    std::vector<ObjMtl> mesh_mtls; std::vector<int> mesh_face_mtl;      // USEMTL extension, resolved after the whole file is read
    std::string line;
    ptd_status rc = PTD_OK;
    while (in.good() && rc == PTD_OK) {
        safeGetline(in, line);
        if (line.empty()) continue;
        std::vector<std::string> tk = tokenize(line);
        if (tk.empty()) continue;
        if (tk[0] == "MATERIAL") {                                                    // scene.cpp:161-196
            if (toki(tk, 1) != (int)sc->materials.size()) continue;                  // id mismatch: block skipped line by line
            ptd_material m;
            memset(&m, 0, sizeof m);
            for (int i = 0; i < 7; ++i) {
                safeGetline(in, line);
                std::vector<std::string> t = tokenize(line);
                if (t.empty()) continue;
                if (t[0] == "RGB") m.color = tok3(t);
                else if (t[0] == "SPECEX") m.specular_exponent = tokf(t, 1);
                else if (t[0] == "SPECRGB") m.specular_color = tok3(t);
                else if (t[0] == "REFL") m.hasReflective = tokf(t, 1);
                else if (t[0] == "REFR") m.hasRefractive = tokf(t, 1);
                else if (t[0] == "REFRIOR") m.indexOfRefraction = tokf(t, 1);
                else if (t[0] == "EMITTANCE") m.emittance = tokf(t, 1);
            }
            sc->materials.push_back(m);
        } else if (tk[0] == "OBJECT") {                                               // scene.cpp:44-100
            if (toki(tk, 1) != (int)sc->geoms.size()) continue;
            ptd_geom g;
            memset(&g, 0, sizeof g);
            g.type = -1;
            safeGetline(in, line);
            if (!line.empty() && in.good()) {
                if (line == "sphere") g.type = PTD_SPHERE;                            // whole-line strcmp, scene.cpp:57-63
                else if (line == "cube") g.type = PTD_CUBE;
            }
            safeGetline(in, line);
            if (!line.empty() && in.good()) g.materialid = toki(tokenize(line), 1);
            safeGetline(in, line);
            while (!line.empty() && in.good()) {
                std::vector<std::string> t = tokenize(line);
                if (!t.empty()) {
                    if (t[0] == "TRANS") g.translation = tok3(t);
                    else if (t[0] == "ROTAT") g.rotation = tok3(t);
                    else if (t[0] == "SCALE") g.scale = tok3(t);
                    else if (t[0] == "VEL") g.vel = tok3(t);
                }
                safeGetline(in, line);
            }
            M4 T = buildTransformationMatrix(g.translation, g.rotation, g.scale);
            M4 I = inverse(T), IT = inverseTranspose(T);
            memcpy(g.transform, T.m, 64); memcpy(g.inverseTransform, I.m, 64); memcpy(g.invTranspose, IT.m, 64);
            sc->geoms.push_back(g);
        } else if (tk[0] == "CAMERA") {                                               // scene.cpp:102-159
            ptd_camera& cam = sc->camera;
            float fovy = 0.f;
            for (int i = 0; i < 5; ++i) {
                safeGetline(in, line);
                std::vector<std::string> t = tokenize(line);
                if (t.empty()) continue;
                if (t[0] == "RES") { cam.res_x = toki(t, 1); cam.res_y = toki(t, 2); }
                else if (t[0] == "FOVY") fovy = tokf(t, 1);
                else if (t[0] == "ITERATIONS") sc->iterations = toki(t, 1);
                else if (t[0] == "DEPTH") sc->trace_depth = toki(t, 1);
                else if (t[0] == "FILE") sc->image_name = t.size() > 1 ? t[1] : "";
            }
            safeGetline(in, line);
            while (!line.empty() && in.good()) {
                std::vector<std::string> t = tokenize(line);
                if (!t.empty()) {
                    if (t[0] == "EYE") cam.position = tok3(t);
                    else if (t[0] == "LOOKAT") cam.lookAt = tok3(t);
                    else if (t[0] == "UP") cam.up = tok3(t);
                }
                safeGetline(in, line);
            }
            if (cam.res_x <= 0 || cam.res_y <= 0) { ptd_set_error("%s: CAMERA block without a valid RES line", path); rc = PTD_ERR_PARSE; break; }
            sc->fovy_deg = fovy;
            ptd_camera_derive(cam, fovy);
            cam.right = norm3(cross3(cam.view, cam.up));      // computed from the still-zero view (scene.cpp:148 before :152) -> NaN; runCuda overwrites it
            cam.view = norm3(sub3(cam.lookAt, cam.position));
            have_camera = true;
        } else if (tk[0] == "MESH") {                                                 // scene.cpp:206-320
            if (toki(tk, 1) != 0) continue;                                           // "Max number of meshes == 1"
            sc->mesh_box.lb = v3(FLT_MAX, FLT_MAX, FLT_MAX);
            sc->mesh_box.ub = v3(FLT_MIN, FLT_MIN, FLT_MIN);                          // numeric_limits<float>::min(), sic (:216-218)
            std::string obj;
            safeGetline(in, line);
            if (!line.empty() && in.good()) { std::vector<std::string> t = tokenize(line); if (t.size() > 1 && t[0] == "PATH") obj = t[1]; }
            int materialid = 0;
            safeGetline(in, line);
            if (!line.empty() && in.good()) materialid = toki(tokenize(line), 1);
            ptd_vec3 tr = v3(0, 0, 0), ro = v3(0, 0, 0), sca = v3(0, 0, 0);
            bool use_mtl = false;
            safeGetline(in, line);
            while (!line.empty() && in.good()) {
                std::vector<std::string> t = tokenize(line);
                if (!t.empty()) {
                    if (t[0] == "TRANS") tr = tok3(t);
                    else if (t[0] == "ROTAT") ro = tok3(t);
                    else if (t[0] == "SCALE") sca = tok3(t);
                    else if (t[0] == "USEMTL") use_mtl = toki(t, 1) != 0;     // extension; the reference ignores unknown keywords here
                }
                safeGetline(in, line);
            }
            M4 T = buildTransformationMatrix(tr, ro, sca);
            // the reference resolves PATH against the process cwd; we also try the scene file's directory
            std::string p1 = obj, p2 = dirname_of(path) + "/" + obj;
            std::ifstream probe(p1.c_str());
            rc = load_obj(*sc, probe.is_open() ? p1 : p2, materialid, T, use_mtl ? &mesh_mtls : nullptr, use_mtl ? &mesh_face_mtl : nullptr);
        }
    }
    if (rc == PTD_OK && !mesh_mtls.empty()) {
        // MTL materials go behind every MATERIAL block of the scene file (their ids are positions in the file, wherever MESH stands)
        const int base = (int)sc->materials.size();
        for (const ObjMtl& m : mesh_mtls) sc->materials.push_back(m.m);
        for (size_t i = 0; i < mesh_face_mtl.size() && i < sc->faces.size(); ++i)
            if (mesh_face_mtl[i] >= 0) sc->faces[i].materialid = base + mesh_face_mtl[i];
    }
    if (rc == PTD_OK && !have_camera) { ptd_set_error("%s: no CAMERA block", path); rc = PTD_ERR_PARSE; }
    if (rc != PTD_OK) { delete sc; return rc; }
    *out = sc;
    return PTD_OK;
}

extern "C" ptd_status ptd_scene_from_arrays(int ngeoms, const ptd_geom* geoms, int nmaterials, const ptd_material* materials,
                                            int nfaces, const ptd_face* faces, const ptd_aabb* mesh_box, const ptd_camera* camera,
                                            int trace_depth, int iterations, ptd_scene** out) {
    if (!out || !camera || ngeoms < 0 || nmaterials < 0 || nfaces < 0 || (ngeoms && !geoms) || (nmaterials && !materials) || (nfaces && !faces))
        PTD_FAIL(PTD_ERR_ARG, "ptd_scene_from_arrays: bad argument");
    if (camera->res_x <= 0 || camera->res_y <= 0) PTD_FAIL(PTD_ERR_ARG, "ptd_scene_from_arrays: camera resolution %dx%d", camera->res_x, camera->res_y);
    ptd_scene* sc = new ptd_scene();
    sc->geoms.assign(geoms, geoms + ngeoms);
    sc->materials.assign(materials, materials + nmaterials);
    sc->faces.assign(faces, faces + nfaces);
    if (mesh_box) sc->mesh_box = *mesh_box; else memset(&sc->mesh_box, 0, sizeof sc->mesh_box);
    sc->camera = *camera;
    sc->fovy_deg = camera->fov_y;
    sc->trace_depth = trace_depth;
    sc->iterations = iterations;
    *out = sc;
    return PTD_OK;
}
extern "C" void ptd_scene_free(ptd_scene* s) { delete s; }
extern "C" ptd_status ptd_scene_counts(const ptd_scene* s, int out[5]) {
    if (!s || !out) PTD_FAIL(PTD_ERR_ARG, "ptd_scene_counts: null argument");
    out[0] = (int)s->geoms.size(); out[1] = (int)s->materials.size(); out[2] = (int)s->faces.size();
    out[3] = s->trace_depth; out[4] = s->iterations;
    return PTD_OK;
}
extern "C" const ptd_geom* ptd_scene_geoms(const ptd_scene* s) { return s && !s->geoms.empty() ? s->geoms.data() : nullptr; }
extern "C" const ptd_material* ptd_scene_materials(const ptd_scene* s) { return s && !s->materials.empty() ? s->materials.data() : nullptr; }
extern "C" const ptd_face* ptd_scene_faces(const ptd_scene* s) { return s && !s->faces.empty() ? s->faces.data() : nullptr; }
extern "C" const ptd_aabb* ptd_scene_mesh_box(const ptd_scene* s) { return s ? &s->mesh_box : nullptr; }
extern "C" ptd_camera* ptd_scene_camera(ptd_scene* s) { return s ? &s->camera : nullptr; }
extern "C" ptd_status ptd_scene_set_resolution(ptd_scene* s, int w, int h) {
    if (!s || w <= 0 || h <= 0) PTD_FAIL(PTD_ERR_ARG, "ptd_scene_set_resolution: bad argument");
    s->camera.res_x = w; s->camera.res_y = h;
    ptd_camera_derive(s->camera, s->fovy_deg);
    return PTD_OK;
}
extern "C" ptd_status ptd_scene_set_depth(ptd_scene* s, int d) {
    if (!s || d < 1) PTD_FAIL(PTD_ERR_ARG, "ptd_scene_set_depth: bad argument");
    s->trace_depth = d;
    return PTD_OK;
}

// main.cpp:66-78
extern "C" ptd_status ptd_camera_orbit_params(const ptd_camera* cam, float* zoom, float* phi, float* theta) {
    if (!cam || !zoom || !phi || !theta) PTD_FAIL(PTD_ERR_ARG, "ptd_camera_orbit_params: null argument");
    ptd_vec3 view = cam->view;
    ptd_vec3 viewXZ = v3(view.x, 0.0f, view.z), viewZY = v3(0.0f, view.y, view.z);
    *phi = acosf(dot3(norm3(viewXZ), v3(0, 0, -1)));
    *theta = acosf(dot3(norm3(viewZY), v3(0, 1, 0)));
    *zoom = len3(sub3(cam->position, cam->lookAt));
    return PTD_OK;
}
// main.cpp:126-138 (cam.right is deliberately NOT normalised, like the reference)
extern "C" ptd_status ptd_camera_orbit(ptd_camera* cam, float zoom, float phi, float theta) {
    if (!cam) PTD_FAIL(PTD_ERR_ARG, "ptd_camera_orbit: null argument");
    ptd_vec3 cp;
    cp.x = zoom * sinf(phi) * sinf(theta);
    cp.y = zoom * cosf(theta);
    cp.z = zoom * cosf(phi) * sinf(theta);
    ptd_vec3 n = norm3(cp);
    cam->view = v3(-n.x, -n.y, -n.z);
    ptd_vec3 v = cam->view, u = v3(0, 1, 0);
    ptd_vec3 r = cross3(v, u);
    cam->up = cross3(r, v);
    cam->right = r;
    cam->position = add3(cp, cam->lookAt);
    return PTD_OK;
}

// ---- BVH build: binned SAH, host side, once per scene ---------------------------------------------------
namespace {
struct Box { float lo[3], hi[3]; };
inline void box_reset(Box& b) { for (int a = 0; a < 3; ++a) { b.lo[a] = FLT_MAX; b.hi[a] = -FLT_MAX; } }
inline void box_grow(Box& b, const float* p) { for (int a = 0; a < 3; ++a) { b.lo[a] = std::min(b.lo[a], p[a]); b.hi[a] = std::max(b.hi[a], p[a]); } }
inline void box_merge(Box& b, const Box& o) { for (int a = 0; a < 3; ++a) { b.lo[a] = std::min(b.lo[a], o.lo[a]); b.hi[a] = std::max(b.hi[a], o.hi[a]); } }
inline float box_area(const Box& b) {
    float d[3] = {b.hi[0] - b.lo[0], b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]};
    if (d[0] < 0) return 0.f;
    return 2.f * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
}
}  // namespace

void ptd_build_bvh(const std::vector<ptd_face>& faces, PtdBvh& out) {
    const int n = (int)faces.size();
    out.nodes.clear(); out.tris.clear(); out.leaves = out.max_leaf = out.max_depth = 0;
    if (n == 0) return;
    std::vector<Box> tb(n);
    std::vector<float> cen((size_t)n * 3);
    Box all; box_reset(all);
    for (int i = 0; i < n; ++i) {
        box_reset(tb[i]);
        for (int v = 0; v < 3; ++v) box_grow(tb[i], &faces[i].v[v].x);
        for (int a = 0; a < 3; ++a) cen[(size_t)i * 3 + a] = 0.5f * (tb[i].lo[a] + tb[i].hi[a]);
        box_merge(all, tb[i]);
    }
    // Conservative padding: a hit point computed in fp32 (and the slab test itself) may be off by a few ulps of the
    // scene extent; every triangle box is grown by `pad` so traversal never culls a face brute force would hit.
    float diag = sqrtf((all.hi[0] - all.lo[0]) * (all.hi[0] - all.lo[0]) + (all.hi[1] - all.lo[1]) * (all.hi[1] - all.lo[1]) +
                       (all.hi[2] - all.lo[2]) * (all.hi[2] - all.lo[2]));
    float maxabs = 0.f;
    for (int a = 0; a < 3; ++a) maxabs = std::max(maxabs, std::max(fabsf(all.lo[a]), fabsf(all.hi[a])));
    const float pad = 3e-5f * std::max(diag, maxabs) + 1e-6f;
    for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) { tb[i].lo[a] -= pad; tb[i].hi[a] += pad; }

    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    struct Task { int node, lo, hi, depth; };
    std::vector<Task> stack;
    out.nodes.reserve((size_t)n);
    out.nodes.push_back(PtdBvhNode());
    stack.push_back(Task{0, 0, n, 1});
    // Build knobs (environment, read per build; the defaults are the configuration every committed measurement used):
    //   PTD_BVH_MAX_LEAF  1..16  ranges of at most this many triangles become leaves                          (4)
    //   PTD_BVH_SWEEP     n      ranges of at most n triangles get an exact sweep-SAH split instead of bins   (0 = never)
    //   PTD_BVH_LEAF_COST x      > 0: SAH termination - a range of <= 16 triangles becomes a leaf when no split beats
    //                            area * count * x (x = cost of a triangle test relative to a node visit)       (0 = off)
    auto env_int = [](const char* k, int def, int lo, int hi) { const char* e = getenv(k); if (!e) return def; const int v = atoi(e); return v < lo || v > hi ? def : v; };
    const int NB = 16, MAX_LEAF = env_int("PTD_BVH_MAX_LEAF", 4, 1, 16), SWEEP = env_int("PTD_BVH_SWEEP", 0, 0, 1 << 20);
    const float LEAF_COST = getenv("PTD_BVH_LEAF_COST") ? (float)atof(getenv("PTD_BVH_LEAF_COST")) : 0.f;
    std::vector<int> order;
    order.reserve(n);
    while (!stack.empty()) {
        Task t = stack.back();
        stack.pop_back();
        Box nb, cb;
        box_reset(nb); box_reset(cb);
        for (int i = t.lo; i < t.hi; ++i) { box_merge(nb, tb[idx[i]]); box_grow(cb, &cen[(size_t)idx[i] * 3]); }
        PtdBvhNode& node = out.nodes[t.node];
        for (int a = 0; a < 3; ++a) { node.bmin[a] = nb.lo[a]; node.bmax[a] = nb.hi[a]; }
        const int cnt = t.hi - t.lo;
        out.max_depth = std::max(out.max_depth, t.depth);
        int split = -1;
        if (cnt > MAX_LEAF || (LEAF_COST > 0.f && cnt > 1)) {
            float best = FLT_MAX;
            int best_axis = -1, best_bin = -1;
            for (int a = 0; a < 3; ++a) {
                float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0.f)) continue;
                Box bb[NB]; int bc[NB];
                for (int b = 0; b < NB; ++b) { box_reset(bb[b]); bc[b] = 0; }
                float k = NB * (1.f - 1e-6f) / ext;
                for (int i = t.lo; i < t.hi; ++i) {
                    int b = std::min(NB - 1, std::max(0, (int)((cen[(size_t)idx[i] * 3 + a] - cb.lo[a]) * k)));
                    box_merge(bb[b], tb[idx[i]]); bc[b]++;
                }
                float la[NB], ra[NB]; int lc[NB], rc[NB];
                Box acc; box_reset(acc); int c = 0;
                for (int b = 0; b < NB; ++b) { box_merge(acc, bb[b]); c += bc[b]; la[b] = box_area(acc); lc[b] = c; }
                box_reset(acc); c = 0;
                for (int b = NB - 1; b >= 0; --b) { box_merge(acc, bb[b]); c += bc[b]; ra[b] = box_area(acc); rc[b] = c; }
                for (int b = 0; b < NB - 1; ++b) {
                    if (lc[b] == 0 || rc[b + 1] == 0) continue;
                    float cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
                    if (cost < best) { best = cost; best_axis = a; best_bin = b; }
                }
            }
            // exact sweep for small ranges: every split position along every axis, centroids sorted
            int sweep_axis = -1, sweep_pos = -1;
            if (cnt <= SWEEP) {
                std::vector<int> tmp(idx.begin() + t.lo, idx.begin() + t.hi);
                std::vector<float> ra(cnt);
                for (int a = 0; a < 3; ++a) {
                    std::sort(tmp.begin(), tmp.end(), [&](int x, int y) { const float cx = cen[(size_t)x * 3 + a], cy = cen[(size_t)y * 3 + a]; return cx < cy || (cx == cy && x < y); });
                    Box acc; box_reset(acc);
                    for (int i = cnt - 1; i > 0; --i) { box_merge(acc, tb[tmp[i]]); ra[i] = box_area(acc); }
                    box_reset(acc);
                    for (int i = 1; i < cnt; ++i) {
                        box_merge(acc, tb[tmp[i - 1]]);
                        const float cost = box_area(acc) * i + ra[i] * (cnt - i);
                        if (cost < best) { best = cost; sweep_axis = a; sweep_pos = i; best_axis = -1; }
                    }
                }
            }
            if (LEAF_COST > 0.f && cnt <= 16 && best < FLT_MAX) {
                // SAH termination: splitting costs one more node visit (area of this node) plus the children's triangle tests
                const float here = box_area(nb);
                if (here * cnt * LEAF_COST <= here + best * LEAF_COST) { best_axis = -1; sweep_axis = -1; }
            }
            if (sweep_axis >= 0) {
                std::sort(idx.begin() + t.lo, idx.begin() + t.hi, [&](int x, int y) { const float cx = cen[(size_t)x * 3 + sweep_axis], cy = cen[(size_t)y * 3 + sweep_axis]; return cx < cy || (cx == cy && x < y); });
                split = t.lo + sweep_pos;
            } else if (best_axis >= 0) {
                float ext = cb.hi[best_axis] - cb.lo[best_axis];
                float k = NB * (1.f - 1e-6f) / ext;
                int* first = idx.data() + t.lo;
                int* last = idx.data() + t.hi;
                int* mid = std::partition(first, last, [&](int i) {
                    int b = std::min(NB - 1, std::max(0, (int)((cen[(size_t)i * 3 + best_axis] - cb.lo[best_axis]) * k)));
                    return b <= best_bin;
                });
                split = (int)(mid - idx.data());
                if (split == t.lo || split == t.hi) split = -1;
            }
            if (split < 0 && cnt > 16) {   // degenerate centroids: median split by index keeps leaves bounded
                split = t.lo + cnt / 2;
            }
        }
        if (split < 0) {
            node.first = (int)order.size();
            node.count = cnt;
            for (int i = t.lo; i < t.hi; ++i) order.push_back(idx[i]);
            std::sort(order.end() - cnt, order.end());    // ascending face index inside a leaf
            out.leaves++;
            out.max_leaf = std::max(out.max_leaf, cnt);
        } else {
            int l = (int)out.nodes.size();
            out.nodes.push_back(PtdBvhNode());
            out.nodes.push_back(PtdBvhNode());
            out.nodes[t.node].first = l;
            out.nodes[t.node].count = 0;
            stack.push_back(Task{l + 1, split, t.hi, t.depth + 1});
            stack.push_back(Task{l, t.lo, split, t.depth + 1});
        }
    }
    // ---- flatten to the 64-byte two-children-per-fetch layout ----
    {
        std::vector<int> wide_index(out.nodes.size(), -1);
        int nw = 0;
        for (size_t i = 0; i < out.nodes.size(); ++i) if (out.nodes[i].count == 0) wide_index[i] = nw++;
        auto child_code = [&](int ni) -> int {
            const PtdBvhNode& c = out.nodes[ni];
            if (c.count == 0) return wide_index[ni];
            return ~((c.first << 4) | (c.count - 1));
        };
        if (nw == 0) {                      // the whole mesh is one leaf: synthetic root with an unreachable second child
            PtdBvhWide w;
            const PtdBvhNode& r = out.nodes[0];
            w.f[0] = r.bmin[0]; w.f[1] = r.bmax[0]; w.f[2] = r.bmin[1]; w.f[3] = r.bmax[1];
            w.f[4] = FLT_MAX; w.f[5] = -FLT_MAX; w.f[6] = FLT_MAX; w.f[7] = -FLT_MAX;
            w.f[8] = r.bmin[2]; w.f[9] = r.bmax[2]; w.f[10] = FLT_MAX; w.f[11] = -FLT_MAX;
            int c0 = child_code(0);
            memcpy(&w.f[12], &c0, 4); memcpy(&w.f[13], &c0, 4); w.f[14] = w.f[15] = 0.f;
            out.wide.push_back(w);
        } else {
            out.wide.resize(nw);
            for (size_t i = 0; i < out.nodes.size(); ++i) {
                if (out.nodes[i].count != 0) continue;
                const int l = out.nodes[i].first, r = l + 1;
                const PtdBvhNode& a = out.nodes[l];
                const PtdBvhNode& b = out.nodes[r];
                PtdBvhWide& w = out.wide[wide_index[i]];
                w.f[0] = a.bmin[0]; w.f[1] = a.bmax[0]; w.f[2] = a.bmin[1]; w.f[3] = a.bmax[1];
                w.f[4] = b.bmin[0]; w.f[5] = b.bmax[0]; w.f[6] = b.bmin[1]; w.f[7] = b.bmax[1];
                w.f[8] = a.bmin[2]; w.f[9] = a.bmax[2]; w.f[10] = b.bmin[2]; w.f[11] = b.bmax[2];
                int c0 = child_code(l), c1 = child_code(r);
                memcpy(&w.f[12], &c0, 4); memcpy(&w.f[13], &c1, 4); w.f[14] = w.f[15] = 0.f;
            }
        }
    }
    // ---- collapse to the 4-wide layout: repeatedly open the child with the largest box until four children (or only leaves) ----
    {
        out.wide4.clear(); out.max_depth4 = 0;
        auto area_of = [&](int ni) { Box b; for (int a = 0; a < 3; ++a) { b.lo[a] = out.nodes[ni].bmin[a]; b.hi[a] = out.nodes[ni].bmax[a]; } return box_area(b); };
        auto leaf_code = [&](int ni) { const PtdBvhNode& c = out.nodes[ni]; return ~((c.first << 4) | (c.count - 1)); };
        struct Job { int bin, slot_node, slot, depth; };      // binary node to turn into a wide node; where its index must be written
        std::vector<Job> jobs;
        auto emit = [&](int bin, int depth) -> int {           // creates the wide node of binary interior node `bin`, queues its interior children
            const int me = (int)out.wide4.size();
            out.wide4.push_back(PtdBvh4());
            out.max_depth4 = std::max(out.max_depth4, depth);
            int kids[4], nk = 0;
            if (out.nodes[bin].count != 0) kids[nk++] = bin;   // the whole mesh is one leaf
            else { kids[nk++] = out.nodes[bin].first; kids[nk++] = out.nodes[bin].first + 1; }
            while (nk < 4) {
                int best = -1; float ba = -1.f;
                for (int k = 0; k < nk; ++k) if (out.nodes[kids[k]].count == 0) { float a = area_of(kids[k]); if (a > ba) { ba = a; best = k; } }
                if (best < 0) break;
                const int open = kids[best];
                kids[best] = out.nodes[open].first; kids[nk++] = out.nodes[open].first + 1;
            }
            PtdBvh4& w = out.wide4[me];
            for (int k = 0; k < 4; ++k) {
                int code = 0;
                if (k < nk) {
                    const PtdBvhNode& c = out.nodes[kids[k]];
                    w.f[k] = c.bmin[0]; w.f[4 + k] = c.bmax[0]; w.f[8 + k] = c.bmin[1]; w.f[12 + k] = c.bmax[1]; w.f[16 + k] = c.bmin[2]; w.f[20 + k] = c.bmax[2];
                    if (c.count != 0) code = leaf_code(kids[k]);
                    else jobs.push_back(Job{kids[k], me, k, depth + 1});
                } else {
                    w.f[k] = w.f[8 + k] = w.f[16 + k] = FLT_MAX; w.f[4 + k] = w.f[12 + k] = w.f[20 + k] = -FLT_MAX;
                    code = ~0;                                  // never reached
                }
                memcpy(&w.f[24 + k], &code, 4);
                w.f[28 + k] = 0.f;
            }
            return me;
        };
        emit(0, 1);
        while (!jobs.empty()) {
            Job j = jobs.back();
            jobs.pop_back();
            const int idx = emit(j.bin, j.depth);
            memcpy(&out.wide4[j.slot_node].f[24 + j.slot], &idx, 4);
        }
    }
    out.tris.resize(order.size());
    for (size_t k = 0; k < order.size(); ++k) {
        const ptd_face& f = faces[order[k]];
        PtdBvhTri& t = out.tris[k];
        t.v0[0] = f.v[0].x; t.v0[1] = f.v[0].y; t.v0[2] = f.v[0].z; t.face = order[k];
        t.v1[0] = f.v[1].x; t.v1[1] = f.v[1].y; t.v1[2] = f.v[1].z; t.material = f.materialid;
        t.v2[0] = f.v[2].x; t.v2[1] = f.v[2].y; t.v2[2] = f.v[2].z; t.pad = 0;
    }
}

// ---- host-side probe of the BVH (no GPU): the traversal of pt_trace restated on the CPU, checked against the brute-force loop ----
namespace {
struct ProbeRng { uint32_t s; float next() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); } };
// Moeller-Trumbore with back-face culling (the semantics of glm::intersectRayTriangle, intersect.inl:37-74): t = ray parameter or -1
inline float probe_tri(const float* v0, const float* v1, const float* v2, const float* o, const float* d) {
    const float e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
    const float p[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
    const float a = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (a < 1.1920929e-7f) return -1.f;
    const float f = 1.0f / a;
    const float sv[3] = {o[0] - v0[0], o[1] - v0[1], o[2] - v0[2]};
    const float bx = f * (sv[0] * p[0] + sv[1] * p[1] + sv[2] * p[2]);
    if (bx < 0.f || bx > 1.f) return -1.f;
    const float q[3] = {sv[1] * e1[2] - sv[2] * e1[1], sv[2] * e1[0] - sv[0] * e1[2], sv[0] * e1[1] - sv[1] * e1[0]};
    const float by = f * (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]);
    if (by < 0.f || by + bx > 1.f) return -1.f;
    return f * (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]);
}
inline void probe_consider(float t, int face, float& t_min, int& best) {     // min t, then lowest face index (pathtrace.cu:259-268)
    if (t > 0.f && (t_min > t || (t_min == t && best >= 0 && face < best))) { t_min = t; best = face; }
}
}  // namespace

// out[0] rays whose BVH hit (face, t) differs from brute force (must be 0)   out[1] interior-node visits per ray
// out[2] triangle tests per ray   out[3] deepest traversal stack   out[4] 4-wide nodes   out[5] leaves
// out[6] mean used children per 4-wide node   out[7] rays that hit something (fraction)
namespace { struct ProbeTrace { std::vector<std::vector<int>> nodes, lines; std::vector<unsigned> bin; int bits = 0; ptd_aabb box; }; static thread_local ProbeTrace* g_probe_trace = nullptr; }

extern "C" ptd_status ptd_bvh_probe(const ptd_scene* sc, int nrays, unsigned seed, int brute_rays, double out[8]) {
    if (!sc || !out || nrays < 1 || brute_rays < 0) PTD_FAIL(PTD_ERR_ARG, "ptd_bvh_probe: bad argument");
    const std::vector<ptd_face>& faces = sc->faces;
    const int nf = (int)faces.size();
    if (nf == 0) PTD_FAIL(PTD_ERR_STATE, "ptd_bvh_probe: the scene has no mesh");
    PtdBvh bvh;
    ptd_build_bvh(faces, bvh);
    PtdBvh8 bvh8;                                                       // PTD_BVH8=1: probe the 8-wide quantised layout instead
    const bool wide8 = getenv("PTD_BVH8") && atoi(getenv("PTD_BVH8")) > 0;
    if (wide8) { ptd_build_bvh8(faces, bvh, bvh8); if (!bvh8.ok) PTD_FAIL(PTD_ERR_UNSUPPORTED, "ptd_bvh_probe: this mesh does not fit the BVH8q encoding"); }
    std::vector<unsigned long long> gstack(4 * (size_t)std::max(bvh8.max_depth, 1) + 8);
    ProbeRng rng{seed * 2654435761u + 12345u};
    long long visits = 0, tests = 0, hits = 0; int max_sp = 0, mismatches = 0;
    std::vector<int> stack(3 * bvh.max_depth4 + 8);
    for (int r = 0; r < nrays; ++r) {
        // a diffuse-bounce-like ray: starts just off a random face, uniform direction in the hemisphere of its geometric normal
        const ptd_face& f = faces[std::min(nf - 1, (int)(rng.next() * nf))];
        float u = rng.next(), v = rng.next();
        if (u + v > 1.f) { u = 1.f - u; v = 1.f - v; }
        ptd_vec3 pnt = add3(f.v[0], add3(muls3(sub3(f.v[1], f.v[0]), u), muls3(sub3(f.v[2], f.v[0]), v)));
        ptd_vec3 n = cross3(sub3(f.v[1], f.v[0]), sub3(f.v[2], f.v[0]));
        if (!(len3(n) > 0.f)) { --r; if (rng.next() < 1e-6f) break; continue; }
        n = norm3(n);
        ptd_vec3 d;
        do { d = v3(2.f * rng.next() - 1.f, 2.f * rng.next() - 1.f, 2.f * rng.next() - 1.f); } while (dot3(d, d) > 1.f || dot3(d, d) < 1e-4f);
        d = norm3(d);
        if (dot3(d, n) < 0.f) d = v3(-d.x, -d.y, -d.z);
        const float o[3] = {pnt.x + 0.01f * d.x, pnt.y + 0.01f * d.y, pnt.z + 0.01f * d.z}, dd[3] = {d.x, d.y, d.z};
        // ---- BVH4 traversal, same order and culling rules as pt_trace ----
        float t_min = FLT_MAX; int best = -1;
        const float ooeps = 1e-30f;
        float id[3], ood[3]; int nearp[3];
        for (int a = 0; a < 3; ++a) { id[a] = 1.0f / (fabsf(dd[a]) > ooeps ? dd[a] : copysignf(ooeps, dd[a])); ood[a] = o[a] * id[a]; nearp[a] = id[a] < 0.f; }
        int sp = 0, node = 0;
        const int SENT = 0x76543210;
        stack[0] = SENT;
        if (g_probe_trace) {                                          // ptd_bvh_probe_order: this ray's bin (same key as ray_bin in ptd_pt.cu)
            ProbeTrace& T = *g_probe_trace;
            T.nodes.emplace_back(); T.lines.emplace_back();
            const int mx = (1 << T.bits) - 1; unsigned c[3];
            const float lo[3] = {T.box.lb.x, T.box.lb.y, T.box.lb.z}, hi[3] = {T.box.ub.x, T.box.ub.y, T.box.ub.z};
            for (int a = 0; a < 3; ++a) c[a] = (unsigned)std::min(std::max((int)((o[a] - lo[a]) / std::max(hi[a] - lo[a], 1e-20f) * (float)(1 << T.bits)), 0), mx);
            unsigned morton = 0;
            for (int b = 0; b < T.bits; ++b) morton |= (((c[0] >> b) & 1u) << (3 * b)) | (((c[1] >> b) & 1u) << (3 * b + 1)) | (((c[2] >> b) & 1u) << (3 * b + 2));
            T.bin.push_back((morton << 3) | (dd[0] < 0.f ? 1u : 0u) | (dd[1] < 0.f ? 2u : 0u) | (dd[2] < 0.f ? 4u : 0u));
        }
        if (wide8) {
            // node group = (first interior child, hit bits << 24 | imask), nearest child = highest hit bit; triangle group = (first triangle, mask)
            const int oct = nearp[0] | (nearp[1] << 1) | (nearp[2] << 2);
            unsigned g_base = 0, g_bits = 0x80000000u | 1u;               // the root: pretend slot 7 ^ oct ... of a parent whose only interior child is node 0
            int gsp = 0; bool root = true;
            for (;;) {
                unsigned t_base = 0, t_bits = 0;
                if (g_bits > 0x00ffffffu) {
                    int child;
                    if (root) { child = 0; g_bits = 0; root = false; }
                    else {
                        const int bit = 31 - __builtin_clz(g_bits);
                        g_bits &= ~(1u << bit);
                        if (g_bits > 0x00ffffffu) { gstack[gsp++] = ((unsigned long long)g_base << 32) | g_bits; max_sp = std::max(max_sp, gsp); }
                        const int slot = (7 - (bit - 24)) ^ oct;
                        child = (int)g_base + __builtin_popcount(g_bits & 0xffu & ((1u << slot) - 1u));
                    }
                    ++visits;
                    if (g_probe_trace) g_probe_trace->nodes.back().push_back(child * 80 / 128);      // the 128-byte line a lane's loads start in
                    const PtdBvh8Node& nd = bvh8.nodes[child];
                    unsigned ih = 0, th = 0;
                    ptd_bvh8_node_hits(nd, o, id, oct, t_min * 1.00001f, &ih, &th);
                    g_base = (unsigned)nd.child_base; g_bits = (ih << 24) | nd.imask;
                    t_base = (unsigned)nd.tri_base & 0x00ffffffu; t_bits = th;
                }
                while (t_bits) {
                    const int k = __builtin_ctz(t_bits);
                    t_bits &= t_bits - 1;
                    ++tests;
                    const PtdBvhTri& t = bvh8.tris[t_base + k];
                    if (g_probe_trace) g_probe_trace->lines.back().push_back((int)((t_base + k) * 48 / 128));
                    probe_consider(probe_tri(t.v0, t.v1, t.v2, o, dd), t.face, t_min, best);
                }
                if (g_bits <= 0x00ffffffu) {
                    if (gsp == 0) break;
                    const unsigned long long e = gstack[--gsp];
                    g_base = (unsigned)(e >> 32); g_bits = (unsigned)e;
                }
            }
            node = SENT;
        }
        while (node != SENT) {
            if (node >= 0) {
                ++visits;
                if (g_probe_trace) g_probe_trace->nodes.back().push_back(node);
                const float* w = bvh.wide4[node].f;
                float dist[4]; int code[4];
                const float tlim = t_min * 1.00001f;
                for (int k = 0; k < 4; ++k) {
                    const float nx = w[(nearp[0] ? 4 : 0) + k], fx = w[(nearp[0] ? 0 : 4) + k], ny = w[8 + (nearp[1] ? 4 : 0) + k], fy = w[8 + (nearp[1] ? 0 : 4) + k];
                    const float nz = w[16 + (nearp[2] ? 4 : 0) + k], fz = w[16 + (nearp[2] ? 0 : 4) + k];
                    const float tn = std::max(std::max(nx * id[0] - ood[0], ny * id[1] - ood[1]), std::max(nz * id[2] - ood[2], 0.0f));
                    const float tf = std::min(std::min(fx * id[0] - ood[0], fy * id[1] - ood[1]), fz * id[2] - ood[2]);
                    dist[k] = (tn <= tf && tn <= tlim) ? tn : FLT_MAX;
                    memcpy(&code[k], &w[24 + k], 4);
                }
                for (int i = 1; i < 4; ++i)                       // insertion sort, ascending entry distance
                    for (int j = i; j > 0 && dist[j] < dist[j - 1]; --j) { std::swap(dist[j], dist[j - 1]); std::swap(code[j], code[j - 1]); }
                for (int k = 3; k >= 1; --k) if (dist[k] < FLT_MAX) stack[++sp] = code[k];
                max_sp = std::max(max_sp, sp);
                node = dist[0] < FLT_MAX ? code[0] : stack[sp--];
            } else {
                const int c = ~node, first = c >> 4, cnt = (c & 15) + 1;
                for (int k = first; k < first + cnt; ++k) {
                    ++tests;
                    if (g_probe_trace) g_probe_trace->lines.back().push_back(k * 48 / 128);
                    const PtdBvhTri& t = bvh.tris[k];
                    probe_consider(probe_tri(t.v0, t.v1, t.v2, o, dd), t.face, t_min, best);
                }
                node = stack[sp--];
            }
        }
        if (best >= 0) ++hits;
        if (r < brute_rays) {
            float bt = FLT_MAX; int bb = -1;
            for (int k = 0; k < nf; ++k) probe_consider(probe_tri(&faces[k].v[0].x, &faces[k].v[1].x, &faces[k].v[2].x, o, dd), k, bt, bb);
            if (bb != best || (bb >= 0 && bt != t_min)) ++mismatches;
        }
    }
    long long used = 0;
    for (const PtdBvh4& w : bvh.wide4) for (int k = 0; k < 4; ++k) if (w.f[k] <= w.f[4 + k]) ++used;
    out[0] = mismatches; out[1] = (double)visits / nrays; out[2] = (double)tests / nrays; out[3] = max_sp; out[4] = (double)bvh.wide4.size();
    out[5] = bvh.leaves; out[6] = bvh.wide4.empty() ? 0.0 : (double)used / bvh.wide4.size(); out[7] = (double)hits / nrays;
    if (wide8) {
        used = 0;
        for (const PtdBvh8Node& n8 : bvh8.nodes) used += __builtin_popcount((unsigned)n8.tri_base >> 24);
        out[4] = (double)bvh8.nodes.size(); out[6] = (double)used / bvh8.nodes.size();
    }
    return PTD_OK;
}

// What PTD_PT_RAY_SORT can buy, estimated on the host: the probe's rays (fully incoherent, like the later bounces) are grouped into
// warps of 32 in arrival order and in (origin cell, direction octant) bin order; lanes step through their node / triangle sequences
// together, and every step costs one L1 wavefront per DISTINCT 128-byte line the active lanes touch (7 loads per node, 3 per
// triangle - the model the ncu capture of pt_trace fits: 115 wavefronts per ray measured).
// out[0], out[1] = wavefronts per ray in arrival / bin order; out[2], out[3] = warp steps per ray (node + triangle trips, i.e.
// issue slots) in arrival / bin order; out[4] = bins in use; out[5] = rays.
extern "C" ptd_status ptd_bvh_probe_order(const ptd_scene* sc, int nrays, unsigned seed, int cell_bits, double out[8]) {
    if (!sc || !out || nrays < 32 || cell_bits < 1 || cell_bits > 5) PTD_FAIL(PTD_ERR_ARG, "ptd_bvh_probe_order: bad argument");
    ProbeTrace T;
    T.bits = cell_bits; T.box = sc->mesh_box;
    g_probe_trace = &T;
    double tmp[8];
    ptd_status rc = ptd_bvh_probe(sc, nrays, seed, 0, tmp);
    g_probe_trace = nullptr;
    if (rc != PTD_OK) return rc;
    const int n = (int)T.nodes.size();
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    auto evaluate = [&](double& wavefronts, double& steps) {
        wavefronts = steps = 0;
        std::vector<int> seen;
        for (int w0 = 0; w0 + 32 <= n; w0 += 32) {
            for (int pass = 0; pass < 2; ++pass) {                     // node sequences, then triangle-line sequences
                size_t longest = 0;
                for (int l = 0; l < 32; ++l) longest = std::max(longest, (pass ? T.lines : T.nodes)[order[w0 + l]].size());
                steps += (double)longest;
                for (size_t k = 0; k < longest; ++k) {
                    seen.clear();
                    for (int l = 0; l < 32; ++l) { const std::vector<int>& q = (pass ? T.lines : T.nodes)[order[w0 + l]]; if (k < q.size()) seen.push_back(q[k]); }
                    std::sort(seen.begin(), seen.end());
                    wavefronts += (double)(std::unique(seen.begin(), seen.end()) - seen.begin()) * (pass ? 3 : 7);
                }
            }
        }
        const double rays = (double)(n / 32 * 32);
        wavefronts /= rays; steps /= rays;
    };
    evaluate(out[0], out[2]);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return T.bin[a] < T.bin[b]; });
    evaluate(out[1], out[3]);
    std::vector<unsigned> bins(T.bin);
    std::sort(bins.begin(), bins.end());
    out[4] = (double)(std::unique(bins.begin(), bins.end()) - bins.begin()); out[5] = n; out[6] = out[7] = 0;
    return PTD_OK;
}

void ptd_geom_bounds(const std::vector<ptd_geom>& geoms, std::vector<ptd_aabb>& out) {
    out.resize(geoms.size());
    for (size_t g = 0; g < geoms.size(); ++g) {
        const float* m = geoms[g].transform;
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        bool finite = true;
        for (int k = 0; k < 8; ++k) {           // the unit cube [-.5,.5]^3 contains the radius-.5 sphere too
            const float x = (k & 1) ? 0.5f : -0.5f, y = (k & 2) ? 0.5f : -0.5f, z = (k & 4) ? 0.5f : -0.5f;
            for (int a = 0; a < 3; ++a) {
                const float v = m[a] * x + m[4 + a] * y + m[8 + a] * z + m[12 + a];
                if (!std::isfinite(v)) finite = false;
                lo[a] = std::min(lo[a], v); hi[a] = std::max(hi[a], v);
            }
        }
        float ext = 0.f;
        for (int a = 0; a < 3; ++a) ext = std::max(ext, std::max(hi[a] - lo[a], std::max(fabsf(lo[a]), fabsf(hi[a]))));
        const float pad = 1e-3f * ext + 1e-4f;   // the exact tests pull hit points back by 1e-4 and round in object space
        for (int a = 0; a < 3; ++a) { lo[a] -= pad; hi[a] += pad; }
        if (!finite) { for (int a = 0; a < 3; ++a) { lo[a] = -FLT_MAX; hi[a] = FLT_MAX; } }   // degenerate transform: never cull
        out[g].lb.x = lo[0]; out[g].lb.y = lo[1]; out[g].lb.z = lo[2];
        out[g].ub.x = hi[0]; out[g].ub.y = hi[1]; out[g].ub.z = hi[2];
    }
}
