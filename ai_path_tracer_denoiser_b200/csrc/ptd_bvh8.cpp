// BVH8q - an 8-wide BVH with 8-bit quantised child boxes (80-byte nodes), collapsed from the binary SAH tree of ptd_build_bvh.
//
// Why: pt_trace sits on the L1 wavefront limit - every lane reads its own 128-byte BVH4 node with seven 16-byte
// loads, 17.6 times per incoherent ray.  Eight children per node cut the visits per ray, and 80-byte nodes need five loads per visit.
// The layout follows the idea of compressed wide BVHs (Ylitie, Karras, Laine 2017): a node stores its own box origin and one
// power-of-two scale per axis; a child box is (origin + qlo * scale, origin + qhi * scale) with qlo rounded down and qhi rounded up, so
// the quantised box always CONTAINS the (already padded) child box - traversal stays conservative and the exact triangle tests and
// tie-breaks of pt_trace decide the result, bit for bit as before.
//
//   bytes  0..11  origin (the node box's lower corner), fp32 x 3
//         12..14  ex, ey, ez: biased exponents, scale = 2^(e - 127), as the fp32 exponent field
//         15      imask: bit s set = slot s holds an interior child
//         16..19  child_base: index of the first interior child; interior children are stored contiguously in slot order
//         20..23  tri_base (low 24 bits): first triangle of this node's leaf children, which are stored contiguously in slot order;
//                 high 8 bits: valid mask (bit s set = slot s holds a child at all)
//         24..31  meta[s]: interior: 0x80 | rank among the interior children; leaf: (count - 1) << 5 | offset from tri_base; empty: 0xff
//         32..79  qlo_x[8] qlo_y[8] qlo_z[8] qhi_x[8] qhi_y[8] qhi_z[8]
// Slot order = traversal order: the child in slot s lies towards the corner (bit 0 of s: +x, bit 1: +y, bit 2: +z) of the node, so
// a ray whose direction signs are `oct` (bit set = negative) meets the slots roughly front to back in ascending (s ^ oct).
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "ptd_internal.h"

namespace {
inline float scale_of(uint8_t e) { uint32_t u = (uint32_t)e << 23; float f; memcpy(&f, &u, 4); return f; }
}

void ptd_build_bvh8(const std::vector<ptd_face>& faces, const PtdBvh& bin, PtdBvh8& out) {
    out.nodes.clear(); out.tris.clear(); out.max_depth = 0; out.max_node_tris = 0; out.ok = false;
    if (bin.nodes.empty()) return;
    bool encodable = true;
    // ---- which binary nodes become wide nodes / leaves: SAH-optimal collapse by dynamic programming over the binary tree (the scheme
    // of the compressed-wide-BVH paper): C[n][i] = cheapest way to represent the subtree of n by at most i children of one wide node.
    const int NB = (int)bin.nodes.size();
    const float CP = getenv("PTD_BVH8_CPRIM") ? (float)atof(getenv("PTD_BVH8_CPRIM")) : 0.4f;      // triangle test relative to a node visit
    const bool greedy = getenv("PTD_BVH8_GREEDY") && atoi(getenv("PTD_BVH8_GREEDY")) > 0;          // the largest-area-first collapse, for comparison
    std::vector<float> C((size_t)NB * 8, 0.f);                        // C[n * 8 + i], i = 1 .. 7
    std::vector<unsigned char> split((size_t)NB * 9, 0), as_leaf(NB, 0);   // split[n * 9 + j]: children given to the left subtree by distribute(n, j)
    std::vector<int> sub_first(NB, 0), sub_tris(NB, 0);
    auto area_n = [&](int ni) {
        const PtdBvhNode& b = bin.nodes[ni];
        const float d[3] = {b.bmax[0] - b.bmin[0], b.bmax[1] - b.bmin[1], b.bmax[2] - b.bmin[2]};
        return d[0] < 0 ? 0.f : 2.f * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
    };
    for (int ni = NB - 1; ni >= 0; --ni) {                              // children have larger indices than their parent
        const PtdBvhNode& b = bin.nodes[ni];
        const float A = area_n(ni);
        if (b.count != 0) {
            sub_first[ni] = b.first; sub_tris[ni] = b.count; as_leaf[ni] = 1;
            for (int i = 1; i <= 7; ++i) C[(size_t)ni * 8 + i] = A * b.count * CP;
            continue;
        }
        const int l = b.first, r = b.first + 1;
        sub_first[ni] = std::min(sub_first[l], sub_first[r]); sub_tris[ni] = sub_tris[l] + sub_tris[r];
        float D[9];
        for (int jn = 2; jn <= 8; ++jn) {
            float best = FLT_MAX; int bk = 1;
            for (int k = 1; k < jn; ++k) {
                if (k > 7 || jn - k > 7) continue;
                const float c = C[(size_t)l * 8 + k] + C[(size_t)r * 8 + (jn - k)];
                if (c < best) { best = c; bk = k; }
            }
            D[jn] = best; split[(size_t)ni * 9 + jn] = (unsigned char)bk;
        }
        const bool contiguous = sub_first[l] + sub_tris[l] == sub_first[r] || sub_first[r] + sub_tris[r] == sub_first[l];
        const float c_leaf = (sub_tris[ni] <= 4 && contiguous && !greedy) ? A * sub_tris[ni] * CP : FLT_MAX;
        const float c_int = D[8] + A;
        as_leaf[ni] = c_leaf <= c_int;
        C[(size_t)ni * 8 + 1] = std::min(c_leaf, c_int);
        for (int i = 2; i <= 7; ++i) C[(size_t)ni * 8 + i] = std::min(D[i], C[(size_t)ni * 8 + i - 1]);
    }
    // children of the wide node made from binary interior node `ni`
    struct Expand { int node, budget; };
    auto children_of = [&](int ni, int* kids) -> int {
        int nk = 0;
        std::vector<Expand> st;
        const PtdBvhNode& b = bin.nodes[ni];
        const int k0 = split[(size_t)ni * 9 + 8];
        st.push_back(Expand{b.first + 1, 8 - k0}); st.push_back(Expand{b.first, k0});
        while (!st.empty()) {
            Expand e = st.back(); st.pop_back();
            const PtdBvhNode& x = bin.nodes[e.node];
            int i = e.budget;
            while (i > 1 && x.count == 0 && C[(size_t)e.node * 8 + i] == C[(size_t)e.node * 8 + i - 1]) --i;     // a smaller forest was as good
            if (i == 1 || x.count != 0) { kids[nk++] = e.node; continue; }
            const int k = split[(size_t)e.node * 9 + i];
            st.push_back(Expand{x.first + 1, i - k}); st.push_back(Expand{x.first, k});
        }
        return nk;
    };
    struct Job { int bin_node, index, depth; };
    std::vector<Job> queue;
    out.nodes.push_back(PtdBvh8Node());
    queue.push_back(Job{0, 0, 1});
    auto area_of = [&](int ni) {
        const PtdBvhNode& b = bin.nodes[ni];
        const float d[3] = {b.bmax[0] - b.bmin[0], b.bmax[1] - b.bmin[1], b.bmax[2] - b.bmin[2]};
        return d[0] < 0 ? 0.f : 2.f * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
    };
    for (size_t qi = 0; qi < queue.size(); ++qi) {
        const Job j = queue[qi];
        out.max_depth = std::max(out.max_depth, j.depth);
        const PtdBvhNode& me = bin.nodes[j.bin_node];
        // ---- children: open the interior child with the largest box until eight (or only leaves) ----
        int kids[8], nk = 0;
        if (me.count != 0 || as_leaf[j.bin_node]) kids[nk++] = j.bin_node;   // the whole mesh is one leaf
        else if (!greedy) nk = children_of(j.bin_node, kids);
        else { kids[nk++] = me.first; kids[nk++] = me.first + 1; }
        while (greedy && nk < 8) {
            int best = -1; float ba = -1.f;
            for (int k = 0; k < nk; ++k) if (bin.nodes[kids[k]].count == 0) { const float a = area_of(kids[k]); if (a > ba) { ba = a; best = k; } }
            if (best < 0) break;
            const int open = kids[best];
            kids[best] = bin.nodes[open].first; kids[nk++] = bin.nodes[open].first + 1;
        }
        // ---- slots: greedy assignment of the children to the octant corners they lie towards ----
        const float cx = 0.5f * (me.bmin[0] + me.bmax[0]), cy = 0.5f * (me.bmin[1] + me.bmax[1]), cz = 0.5f * (me.bmin[2] + me.bmax[2]);
        int slot_of[8], kid_in[8];
        for (int s = 0; s < 8; ++s) kid_in[s] = -1;
        for (int k = 0; k < nk; ++k) slot_of[k] = -1;
        for (int round = 0; round < nk; ++round) {
            float bs = -FLT_MAX; int bk = -1, bslot = -1;
            for (int k = 0; k < nk; ++k) {
                if (slot_of[k] >= 0) continue;
                const PtdBvhNode& c = bin.nodes[kids[k]];
                const float dx = 0.5f * (c.bmin[0] + c.bmax[0]) - cx, dy = 0.5f * (c.bmin[1] + c.bmax[1]) - cy, dz = 0.5f * (c.bmin[2] + c.bmax[2]) - cz;
                for (int s = 0; s < 8; ++s) {
                    if (kid_in[s] >= 0) continue;
                    const float score = ((s & 1) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 4) ? dz : -dz);
                    if (score > bs) { bs = score; bk = k; bslot = s; }
                }
            }
            slot_of[bk] = bslot; kid_in[bslot] = kids[bk];
        }
        // ---- the node record ----
        PtdBvh8Node n;
        memset(&n, 0, sizeof n);
        n.origin[0] = me.bmin[0]; n.origin[1] = me.bmin[1]; n.origin[2] = me.bmin[2];
        for (int a = 0; a < 3; ++a) {
            const double ext = (double)me.bmax[a] - (double)me.bmin[a];
            int e = -100;                                              // smallest scale used: keeps q * scale * (1/d) a normal number
            if (ext > 0) e = std::max(-100, (int)std::ceil(std::log2(ext / 255.0)));
            while (std::ldexp(255.0, e) < ext) ++e;
            n.e[a] = (uint8_t)(e + 127);
        }
        int n_interior = 0, n_tris = 0;
        auto is_leaf = [&](int ni) { return bin.nodes[ni].count != 0 || as_leaf[ni]; };
        for (int s = 0; s < 8; ++s) if (kid_in[s] >= 0 && !is_leaf(kid_in[s])) ++n_interior;
        n.child_base = (int)out.nodes.size();
        out.nodes.resize(out.nodes.size() + n_interior);              // the interior children, contiguous, in slot order
        const int tri_base = (int)out.tris.size();
        unsigned valid = 0; int rank = 0;
        for (int s = 0; s < 8; ++s) {
            for (int a = 0; a < 3; ++a) { n.qlo[a][s] = 255; n.qhi[a][s] = 0; }
            n.meta[s] = 0xff;
            if (kid_in[s] < 0) continue;
            valid |= 1u << s;
            const PtdBvhNode& c = bin.nodes[kid_in[s]];
            for (int a = 0; a < 3; ++a) {
                const double sc = std::ldexp(1.0, (int)n.e[a] - 127), o = n.origin[a];
                int lo = (int)std::floor(((double)c.bmin[a] - o) / sc), hi = (int)std::ceil(((double)c.bmax[a] - o) / sc);
                lo = std::min(std::max(lo, 0), 255); hi = std::min(std::max(hi, 0), 255);
                while (lo > 0 && o + lo * sc > (double)c.bmin[a]) --lo;                  // conservative whatever the rounding above did
                while (hi < 255 && o + hi * sc < (double)c.bmax[a]) ++hi;
                n.qlo[a][s] = (uint8_t)lo; n.qhi[a][s] = (uint8_t)hi;
            }
            if (!is_leaf(kid_in[s])) {
                n.imask |= (uint8_t)(1u << s);
                n.meta[s] = (uint8_t)(0x80 | rank);
                queue.push_back(Job{kid_in[s], n.child_base + rank, j.depth + 1});
                ++rank;
            } else {
                const int cnt = sub_tris[kid_in[s]], first = sub_first[kid_in[s]];       // a binary leaf, or a small subtree merged into one leaf
                if (cnt > 4 || n_tris + cnt > 32) encodable = false;                      // 2 count bits, 5 offset bits, 32-bit triangle mask
                n.meta[s] = (uint8_t)((((cnt - 1) & 3) << 5) | (n_tris & 31));
                for (int k = 0; k < cnt; ++k) out.tris.push_back(bin.tris[first + k]);
                n_tris += cnt;
            }
        }
        out.max_node_tris = std::max(out.max_node_tris, n_tris);
        n.tri_base = tri_base | (int)(valid << 24);
        out.nodes[j.index] = n;
    }
    out.ok = encodable && out.tris.size() < (1u << 24);                  // else the caller keeps the 4-wide layout
    (void)faces;
}

// ---- the traversal pt_trace_bvh8 runs, restated on the host (ptd_bvh_probe with PTD_BVH8=1) -------------------------------------------
// Returns the hit child bits of one node: interior hits as priority bits (bit 7 - (slot ^ oct) -> nearest first), leaf hits as
// bits of the node's triangle list.
void ptd_bvh8_node_hits(const PtdBvh8Node& n, const float o[3], const float idir[3], int oct, float tlim, unsigned* interior_hits, unsigned* tri_hits) {
    float a[3], b[3];
    for (int k = 0; k < 3; ++k) { a[k] = scale_of(n.e[k]) * idir[k]; b[k] = (n.origin[k] - o[k]) * idir[k]; }
    const unsigned valid = (unsigned)n.tri_base >> 24;
    unsigned ih = 0, th = 0;
    for (int s = 0; s < 8; ++s) {
        if (!((valid >> s) & 1u)) continue;
        float tn = 0.0f, tf = FLT_MAX;
        for (int k = 0; k < 3; ++k) {
            const float lo = (float)n.qlo[k][s] * a[k] + b[k], hi = (float)n.qhi[k][s] * a[k] + b[k];
            tn = std::max(tn, (oct >> k) & 1 ? hi : lo);
            tf = std::min(tf, (oct >> k) & 1 ? lo : hi);
        }
        if (!(tn <= tf && tn <= tlim)) continue;
        const unsigned m = n.meta[s];
        if (m & 0x80u) ih |= 1u << (7 - (s ^ oct));
        else th |= ((2u << (m >> 5)) - 1u) << (m & 31u);
    }
    *interior_hits = ih; *tri_hits = th;
}
