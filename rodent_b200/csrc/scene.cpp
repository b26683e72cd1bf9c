// Runtime scene construction (host side): what the reference does at build time in
// its scene compiler src/driver/converter.cpp, done when a scene is loaded.
//
//   OBJ / MTL parsing                 src/driver/obj.cpp:104-255, 257-371
//   material clean-up, de-duplication src/driver/converter.cpp:440-557
//   triangle mesh (dedup, normals)    src/driver/obj.cpp:412-509
//   lights                            src/driver/converter.cpp:770-818
//   material -> BSDF rules            src/driver/converter.cpp:857-913
//   BVH8 / Tri4 emission              src/driver/converter.cpp:160-260 (node / leaf writers)
//
// The BVH itself is built with a binned-SAH top-down build followed by a collapse to
// arity 8 (the reference uses an SBVH builder, src/driver/bvh.h; tree quality only
// affects speed, not results).  The Impala code generation of the reference is replaced
// by the RodentMaterial / RodentLight tables of include/rodent_b200.h.
#include "scene.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <sstream>
#include <unordered_map>

namespace rb200 {
namespace {

struct F3 {
    float x = 0, y = 0, z = 0;
    bool operator==(const F3& o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(const F3& o) const { return !(*this == o); }
};
F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
F3 operator*(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
F3 cross(F3 a, F3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
float length(F3 a) { return std::sqrt(dot(a, a)); }
F3 normalize(F3 a) { return a * (1.0f / length(a)); }      // src/driver/float3.h:139-141
F3 fmin3(F3 a, F3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
F3 fmax3(F3 a, F3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }

struct ObjIndex { int v = 0, t = 0, n = 0; };
struct ObjFace { std::vector<ObjIndex> idx; int material = 0; };
struct ObjMaterial {                      // src/driver/obj.h:31-48
    F3 ka, kd, ks, ke, tf;
    float ns = 0, ni = 0, tr = 0, d = 0;
    int illum = 0;
    std::string map_ka, map_kd, map_ks, map_ke, map_bump, map_d;
};
struct ObjFile {
    std::vector<std::vector<ObjFace>> objects;   // faces per `o` object (groups do not matter downstream)
    std::vector<F3> vertices, normals;
    std::vector<std::array<float, 2>> texcoords;
    std::vector<std::string> materials, mtl_libs;
};

void warn(const std::string& m) { std::cerr << "rodent_b200: warning: " << m << std::endl; }
bool fail(const std::string& m) { std::cerr << "rodent_b200: error: " << m << std::endl; return false; }

const char* skip_ws(const char* p) { while (*p && std::isspace((unsigned char)*p)) p++; return p; }
std::string first_word(const char* p) {
    p = skip_ws(p);
    const char* e = p;
    while (*e && !std::isspace((unsigned char)*e)) e++;
    return std::string(p, e);
}
std::string rest_of_line(const char* p) {
    std::string s = skip_ws(p);
    while (!s.empty() && std::isspace((unsigned char)s.back())) s.pop_back();
    return s;
}
bool keyword(const char* p, const char* kw) {
    const size_t n = std::strlen(kw);
    return !std::strncmp(p, kw, n) && std::isspace((unsigned char)p[n]);
}
void read_floats(const char* p, float* out, int n) {
    char* e;
    for (int i = 0; i < n; i++) { out[i] = std::strtof(p, &e); p = e; }
}

// v, v/t, v//n, v/t/n with negative (relative) indices: obj.cpp:69-100
bool read_index(const char*& p, ObjIndex& idx) {
    p = skip_ws(p);
    if (!std::isdigit((unsigned char)*p) && *p != '-') return false;
    char* e;
    idx = ObjIndex{};
    idx.v = int(std::strtol(p, &e, 10)); p = skip_ws(e);
    if (*p == '/') {
        p++;
        if (*p != '/') { idx.t = int(std::strtol(p, &e, 10)); p = e; }
        p = skip_ws(p);
        if (*p == '/') { p++; idx.n = int(std::strtol(p, &e, 10)); p = e; }
    }
    return true;
}

bool parse_obj(const std::string& path, ObjFile& file) {
    std::ifstream in(path);
    if (!in) return fail("cannot open OBJ file '" + path + "'");
    file.objects.emplace_back();
    file.materials.emplace_back("");                       // material 0 = the dummy material
    file.vertices.emplace_back(); file.normals.emplace_back(); file.texcoords.push_back({0, 0});   // 1-based indices
    int cur_mtl = 0, line_no = 0, errors = 0;
    std::string line;
    while (std::getline(in, line)) {
        line_no++;
        const char* p = skip_ws(line.c_str());
        if (!*p || *p == '#') continue;
        if (p[0] == 'v' && std::isspace((unsigned char)p[1])) {
            float v[3]; read_floats(p + 1, v, 3); file.vertices.push_back({v[0], v[1], v[2]});
        } else if (p[0] == 'v' && p[1] == 'n') {
            float v[3]; read_floats(p + 2, v, 3); file.normals.push_back({v[0], v[1], v[2]});
        } else if (p[0] == 'v' && p[1] == 't') {
            float v[2]; read_floats(p + 2, v, 2); file.texcoords.push_back({v[0], v[1]});
        } else if (p[0] == 'f' && std::isspace((unsigned char)p[1])) {
            ObjFace f; f.material = cur_mtl;
            const char* q = p + 2;
            ObjIndex idx;
            while (read_index(q, idx)) f.idx.push_back(idx);
            bool ok = f.idx.size() >= 3;
            for (auto& i : f.idx) {
                if (i.v < 0) i.v += int(file.vertices.size());
                if (i.t < 0) i.t += int(file.texcoords.size());
                if (i.n < 0) i.n += int(file.normals.size());
                ok = ok && i.v > 0 && i.t >= 0 && i.n >= 0 && i.v < int(file.vertices.size()) &&
                     i.t < int(file.texcoords.size()) && i.n < int(file.normals.size());
            }
            if (ok) file.objects.back().push_back(f);
            else { fail("invalid face (line " + std::to_string(line_no) + ")"); errors++; }
        } else if (p[0] == 'g' && std::isspace((unsigned char)p[1])) {
        } else if (p[0] == 'o' && std::isspace((unsigned char)p[1])) {
            file.objects.emplace_back();
        } else if (keyword(p, "usemtl")) {
            const std::string name = first_word(p + 6);
            cur_mtl = int(std::find(file.materials.begin(), file.materials.end(), name) - file.materials.begin());
            if (cur_mtl == int(file.materials.size())) file.materials.push_back(name);
        } else if (keyword(p, "mtllib")) {
            file.mtl_libs.push_back(first_word(p + 6));
        } else if (p[0] == 's' && std::isspace((unsigned char)p[1])) {
        } else { fail("unknown OBJ command '" + std::string(p) + "' (line " + std::to_string(line_no) + ")"); errors++; }
    }
    return errors == 0;
}

bool parse_mtl(const std::string& path, std::map<std::string, ObjMaterial>& lib) {
    std::ifstream in(path);
    if (!in) return fail("cannot open MTL file '" + path + "'");
    std::string line, name;
    int errors = 0;
    while (std::getline(in, line)) {
        const char* p = skip_ws(line.c_str());
        if (!*p || *p == '#') continue;
        if (keyword(p, "newmtl")) {
            name = first_word(p + 6);
            if (lib.count(name)) { fail("material redefinition for '" + name + "'"); errors++; }
            lib[name];
            continue;
        }
        ObjMaterial& m = lib[name];
        if (p[0] == 'K' && std::isspace((unsigned char)p[2]) && std::strchr("adse", p[1])) {
            F3& c = p[1] == 'a' ? m.ka : p[1] == 'd' ? m.kd : p[1] == 's' ? m.ks : m.ke;
            float v[3]; read_floats(p + 3, v, 3); c = {v[0], v[1], v[2]};
        } else if (keyword(p, "Ns")) read_floats(p + 3, &m.ns, 1);
        else if (keyword(p, "Ni")) read_floats(p + 3, &m.ni, 1);
        else if (keyword(p, "Tf")) { float v[3]; read_floats(p + 3, v, 3); m.tf = {v[0], v[1], v[2]}; }
        else if (keyword(p, "Tr")) read_floats(p + 3, &m.tr, 1);
        else if (p[0] == 'd' && std::isspace((unsigned char)p[1])) read_floats(p + 2, &m.d, 1);
        else if (keyword(p, "illum")) { float v; read_floats(p + 6, &v, 1); m.illum = int(v); }
        else if (keyword(p, "map_Ka")) m.map_ka = rest_of_line(p + 6);
        else if (keyword(p, "map_Kd")) m.map_kd = rest_of_line(p + 6);
        else if (keyword(p, "map_Ks")) m.map_ks = rest_of_line(p + 6);
        else if (keyword(p, "map_Ke")) m.map_ke = rest_of_line(p + 6);
        else if (keyword(p, "map_bump")) m.map_bump = rest_of_line(p + 8);
        else if (keyword(p, "bump")) m.map_bump = rest_of_line(p + 4);
        else if (keyword(p, "map_d")) m.map_d = rest_of_line(p + 5);
        else if (p[0] == 'K' || p[0] == 'N' || p[0] == 'T') { fail("invalid MTL command '" + std::string(p) + "'"); errors++; }
        else warn("unknown MTL command '" + std::string(p) + "'");
    }
    return errors == 0;
}

// converter.cpp:440-458 (tr, d, map_ka, map_bump, map_d are ignored)
bool same_material(const ObjMaterial& a, const ObjMaterial& b) {
    return a.ka == b.ka && a.kd == b.kd && a.ks == b.ks && a.ke == b.ke && a.ns == b.ns && a.ni == b.ni && a.tf == b.tf &&
           a.illum == b.illum && a.map_kd == b.map_kd && a.map_ks == b.map_ks && a.map_ke == b.map_ke;
}
// converter.cpp:460-465
bool is_simple(const ObjMaterial& m) {
    const F3 zero;
    return m.illum != 5 && m.illum != 7 && m.ke == zero && m.map_ke.empty() && m.map_kd.empty() && m.map_ks.empty() &&
           (m.kd != zero || m.ks != zero);
}

// cleanup_obj, converter.cpp:467-557: dummy material, missing -> dummy, identical -> first, unused removed,
// "complex" materials before "simple" ones.
void cleanup_materials(ObjFile& obj, std::map<std::string, ObjMaterial>& lib) {
    ObjMaterial& dummy = lib[""];
    dummy = ObjMaterial{};
    dummy.kd = {0.0f, 1.0f, 1.0f}; dummy.ns = 1.0f; dummy.ni = 1.0f; dummy.tr = 1.0f; dummy.d = 1.0f; dummy.illum = 2;
    for (auto& name : obj.materials)
        if (!name.empty() && !lib.count(name)) { warn("missing material definition for '" + name + "', replaced by the dummy material"); name = ""; }
    std::unordered_map<std::string, std::string> remap;
    for (size_t i = 0; i < obj.materials.size(); i++) {
        if (remap.count(obj.materials[i])) continue;
        for (size_t j = i + 1; j < obj.materials.size(); j++)
            if (same_material(lib[obj.materials[i]], lib[obj.materials[j]])) remap.emplace(obj.materials[j], obj.materials[i]);
    }
    auto canonical = [&](std::string n) { auto it = remap.find(n); return it == remap.end() ? n : it->second; };
    std::vector<std::string> used;
    for (auto& faces : obj.objects)
        for (auto& f : faces) {
            const std::string n = canonical(obj.materials[f.material]);
            if (std::find(used.begin(), used.end(), n) == used.end()) used.push_back(n);
        }
    std::vector<std::string> kept;
    for (auto& n : obj.materials)
        if (std::find(used.begin(), used.end(), n) != used.end() && std::find(kept.begin(), kept.end(), n) == kept.end()) kept.push_back(n);
    std::stable_partition(kept.begin(), kept.end(), [&](const std::string& n) { return !is_simple(lib[n]); });
    std::vector<int> id_remap;
    for (auto& n : obj.materials) id_remap.push_back(int(std::find(kept.begin(), kept.end(), canonical(n)) - kept.begin()));
    for (auto& faces : obj.objects)
        for (auto& f : faces) f.material = id_remap[f.material];
    obj.materials = kept;
}

// converter.cpp:870-913; map_kd / map_ks are filled in by the caller once the images are loaded
RodentMaterial make_material(const ObjMaterial& m) {
    RodentMaterial r{};
    const F3 zero;
    r.ns = m.ns; r.ni = m.ni;
    r.kd[0] = m.kd.x; r.kd[1] = m.kd.y; r.kd[2] = m.kd.z;
    r.ks[0] = m.ks.x; r.ks[1] = m.ks.y; r.ks[2] = m.ks.z;
    r.tf[0] = m.tf.x; r.tf[1] = m.tf.y; r.tf[2] = m.tf.z;
    r.ke[0] = m.ke.x; r.ke[1] = m.ke.y; r.ke[2] = m.ke.z;
    r.is_emissive = (m.ke != zero || !m.map_ke.empty()) ? 1 : 0;
    if (m.illum == 5) r.bsdf = RODENT_BSDF_MIRROR;
    else if (m.illum == 7) r.bsdf = RODENT_BSDF_GLASS;
    else {
        const bool diffuse = m.kd != zero || !m.map_kd.empty(), specular = m.ks != zero || !m.map_ks.empty();
        if (diffuse && specular) {
            // color_luminance, src/core/color.impala:33-35
            const float lum_ks = m.ks.x * 0.2126f + m.ks.y * 0.7152f + m.ks.z * 0.0722f;
            const float lum_kd = m.kd.x * 0.2126f + m.kd.y * 0.7152f + m.kd.z * 0.0722f;
            r.mix_k = (lum_ks + lum_kd == 0.0f) ? 0.0f : lum_ks / (lum_ks + lum_kd);
            r.bsdf = RODENT_BSDF_MIX;
        } else r.bsdf = diffuse ? RODENT_BSDF_DIFFUSE : specular ? RODENT_BSDF_PHONG : RODENT_BSDF_BLACK;
    }
    return r;
}

// ---- BVH8 / Tri4 build ---------------------------------------------------------------
struct Box {
    F3 lo{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    F3 hi{-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
    void grow(F3 p) { lo = fmin3(lo, p); hi = fmax3(hi, p); }
    void grow(const Box& b) { lo = fmin3(lo, b.lo); hi = fmax3(hi, b.hi); }
    float half_area() const { const F3 e = hi - lo; return std::max(e.x, 0.f) * (std::max(e.y, 0.f) + std::max(e.z, 0.f)) + std::max(e.y, 0.f) * std::max(e.z, 0.f); }
};
struct Bvh2Node { Box box; int left = -1, right = -1, first = 0, count = 0; };

// NodeT: Node8 or Node4 -- the binary tree is the same, the collapse stops at the node's arity.
template <typename NodeT>
struct Builder {
    const std::vector<F3>& v0; const std::vector<F3>& v1; const std::vector<F3>& v2;
    const std::vector<int>& geom;
    std::vector<Box> boxes; std::vector<F3> centers; std::vector<int> order;
    std::vector<Bvh2Node> n2;
    std::vector<NodeT>& nodes; std::vector<Tri4>& tris;
    static constexpr int kArity = int(sizeof(NodeT::child) / sizeof(int32_t));

    // ---- the binary tree: a split BVH (Stich, Friedrich, Dietrich: "Spatial Splits in Bounding Volume Hierarchies",
    // HPG 2009), the algorithm of the reference's builder (src/driver/bvh.h:102-246) in this repository's own form.
    // A node owns a list of REFERENCES (triangle id + the part of its bounds inside the node).  Candidates per node:
    //   * object split: references sorted by centroid along each axis, full sweep of n * area(left) + n * area(right);
    //   * spatial split, tried when the object split's children overlap by more than kAlpha of the root's area: the
    //     node's extent cut into kSpatialBins slabs per axis, triangles clipped against the slab planes, sweep over the
    //     planes; a reference straddling the chosen plane goes to both sides (clipped) unless keeping it whole on one
    //     side is cheaper ("unsplitting");
    // and the node becomes a leaf when the best split does not pay: cost(split) + area >= n * area
    // (CostFn of converter.cpp:118-128: leaf_cost = count * area, traversal_cost = area), or at kLeafMin references.
    // Leaves append their triangle ids to `order` (an id can then appear in several leaves).
    struct Ref { int id; Box box; };
    static constexpr int kSpatialBins = 64, kLeafMin = 2;
    static constexpr float kAlpha = 1e-5f;
    float root_area = 0.0f;
    int spatial_splits = 0;

    static float axis_of(const F3& v, int axis) { return (&v.x)[axis]; }
    static float& axis_of(F3& v, int axis) { return (&v.x)[axis]; }
    static Box intersect(const Box& a, const Box& b) { Box r; r.lo = fmax3(a.lo, b.lo); r.hi = fmin3(a.hi, b.hi); return r; }
    static float ref_center(const Ref& r, int axis) { return 0.5f * (axis_of(r.box.lo, axis) + axis_of(r.box.hi, axis)); }

    // Bounds of the two parts of triangle `id` on either side of the plane `axis = pos`.
    void clip_triangle(int id, int axis, float pos, Box& left, Box& right) const {
        const F3 v[3] = {v0[id], v1[id], v2[id]};
        left = Box(); right = Box();
        for (int k = 0; k < 3; k++) {
            const F3 a = v[k], b = v[(k + 1) % 3];
            const float pa = axis_of(a, axis), pb = axis_of(b, axis);
            if (pa <= pos) left.grow(a);
            if (pa >= pos) right.grow(a);
            if ((pa < pos && pb > pos) || (pa > pos && pb < pos)) {
                const float t = (pos - pa) / (pb - pa);
                F3 x = a + (b - a) * t;
                axis_of(x, axis) = pos;
                left.grow(x); right.grow(x);
            }
        }
    }

    struct ObjectSplit { float cost = std::numeric_limits<float>::max(); int axis = -1, left_count = 0; Box left, right; };
    struct SpatialSplit { float cost = std::numeric_limits<float>::max(); int axis = -1; float pos = 0; };

    void find_object_split(std::vector<Ref>& refs, ObjectSplit& best, std::vector<Box>& scratch) {
        const int n = int(refs.size());
        scratch.resize(n);
        for (int axis = 0; axis < 3; axis++) {
            std::sort(refs.begin(), refs.end(), [axis](const Ref& a, const Ref& b) {
                const float ca = ref_center(a, axis), cb = ref_center(b, axis);
                return ca < cb || (ca == cb && a.id < b.id);
            });
            Box acc;
            for (int i = n - 1; i > 0; i--) { acc.grow(refs[i].box); scratch[i] = acc; }
            acc = Box();
            for (int i = 0; i < n - 1; i++) {
                acc.grow(refs[i].box);
                const float c = float(i + 1) * acc.half_area() + float(n - i - 1) * scratch[i + 1].half_area();
                if (c < best.cost) { best.cost = c; best.axis = axis; best.left_count = i + 1; best.left = acc; best.right = scratch[i + 1]; }
            }
        }
    }

    void find_spatial_split(const std::vector<Ref>& refs, const Box& box, SpatialSplit& best) {
        const int n = int(refs.size());
        for (int axis = 0; axis < 3; axis++) {
            const float lo = axis_of(box.lo, axis), hi = axis_of(box.hi, axis);
            if (!(hi > lo)) continue;
            const float width = (hi - lo) / kSpatialBins, inv = 1.0f / width;
            Box bins[kSpatialBins]; int enter[kSpatialBins] = {}, leave[kSpatialBins] = {};
            auto bin_of = [&](float x) { return std::max(0, std::min(kSpatialBins - 1, int((x - lo) * inv))); };
            for (const Ref& r : refs) {
                const int first = bin_of(axis_of(r.box.lo, axis)), last = bin_of(axis_of(r.box.hi, axis));
                Box rest = r.box;
                for (int b = first; b < last; b++) {                  // chop the reference at every slab plane it crosses
                    Box l, rr;
                    clip_triangle(r.id, axis, lo + float(b + 1) * width, l, rr);
                    bins[b].grow(intersect(l, rest));
                    rest = intersect(rest, rr);
                }
                bins[last].grow(rest);
                enter[first]++; leave[last]++;
            }
            Box right_acc[kSpatialBins]; Box acc;
            for (int b = kSpatialBins - 1; b > 0; b--) { acc.grow(bins[b]); right_acc[b] = acc; }
            acc = Box();
            int nl = 0, nr = n;
            for (int b = 0; b < kSpatialBins - 1; b++) {
                acc.grow(bins[b]); nl += enter[b]; nr -= leave[b];
                if (nl == 0 || nr == 0) continue;
                const float c = float(nl) * acc.half_area() + float(nr) * right_acc[b + 1].half_area();
                if (c < best.cost) { best.cost = c; best.axis = axis; best.pos = lo + float(b + 1) * width; }
            }
        }
    }

    // Distributes the references over the two sides of the chosen plane; returns false when one side stays empty.
    bool apply_spatial_split(std::vector<Ref>& refs, const SpatialSplit& sp, std::vector<Ref>& left, std::vector<Ref>& right, Box& lbox, Box& rbox) {
        std::vector<Ref> straddling;
        lbox = Box(); rbox = Box();
        for (const Ref& r : refs) {
            if (axis_of(r.box.hi, sp.axis) <= sp.pos) { left.push_back(r); lbox.grow(r.box); }
            else if (axis_of(r.box.lo, sp.axis) >= sp.pos) { right.push_back(r); rbox.grow(r.box); }
            else straddling.push_back(r);
        }
        for (const Ref& r : straddling) {
            Box l, rr;
            clip_triangle(r.id, sp.axis, sp.pos, l, rr);
            l = intersect(l, r.box); rr = intersect(rr, r.box);
            Box l_whole = lbox, r_whole = rbox, l_part = lbox, r_part = rbox;
            l_whole.grow(r.box); r_whole.grow(r.box); l_part.grow(l); r_part.grow(rr);
            const float nl = float(left.size()), nr = float(right.size());
            const float to_left = (nl + 1) * l_whole.half_area() + nr * rbox.half_area();
            const float to_right = nl * lbox.half_area() + (nr + 1) * r_whole.half_area();
            const float both = (nl + 1) * l_part.half_area() + (nr + 1) * r_part.half_area();
            if (to_left <= to_right && to_left <= both) { left.push_back(r); lbox = l_whole; }
            else if (to_right <= both) { right.push_back(r); rbox = r_whole; }
            else { left.push_back(Ref{r.id, l}); right.push_back(Ref{r.id, rr}); lbox = l_part; rbox = r_part; }
        }
        return !left.empty() && !right.empty() && left.size() < refs.size() && right.size() < refs.size();   // both sides must shrink
    }

    int make_leaf2(int id, const std::vector<Ref>& refs) {
        n2[id].first = int(order.size());
        n2[id].count = int(refs.size());
        for (const Ref& r : refs) order.push_back(r.id);
        return id;
    }

    int build2(std::vector<Ref>& refs, const Box& box, int depth) {
        const int id = int(n2.size());
        n2.emplace_back();
        n2[id].box = box;
        const int n = int(refs.size());
        if (n <= kLeafMin || depth > 60) return make_leaf2(id, refs);
        std::vector<Box> scratch;
        ObjectSplit os;
        find_object_split(refs, os, scratch);
        SpatialSplit ss;
        if (use_spatial_splits && os.axis >= 0 && intersect(os.left, os.right).half_area() > kAlpha * root_area) find_spatial_split(refs, box, ss);
        const float best = std::min(os.cost, ss.cost);
        // no split pays for itself: a leaf -- unless it would be a long one (more than kLeafMax references cost several
        // Tri4 packets every time a ray enters it), which is split anyway
        const bool split_pays = os.axis >= 0 && best + box.half_area() < float(n) * box.half_area();
        if (!split_pays && n <= kLeafMax) return make_leaf2(id, refs);
        std::vector<Ref> left, right; Box lbox, rbox;
        bool done = false;
        if (ss.cost < os.cost) {
            done = apply_spatial_split(refs, ss, left, right, lbox, rbox);
            if (done) spatial_splits++;
            else { left.clear(); right.clear(); }
        }
        if (!done) {
            if (os.axis < 0) {                      // identical centroids everywhere: halve the list
                os.left_count = n / 2;
                for (int i = 0; i < n; i++) (i < os.left_count ? os.left : os.right).grow(refs[i].box);
            } else if (os.axis != 2) {              // the list is in z order after the sweeps
                const int axis = os.axis;
                std::sort(refs.begin(), refs.end(), [axis](const Ref& a, const Ref& b) {
                    const float ca = ref_center(a, axis), cb = ref_center(b, axis);
                    return ca < cb || (ca == cb && a.id < b.id);
                });
            }
            left.assign(refs.begin(), refs.begin() + os.left_count);
            right.assign(refs.begin() + os.left_count, refs.end());
            lbox = os.left; rbox = os.right;
        }
        std::vector<Ref>().swap(refs);              // the parent's list is not needed below
        const int l = build2(left, lbox, depth + 1);
        const int r = build2(right, rbox, depth + 1);
        n2[id].left = l; n2[id].right = r;
        return id;
    }
    bool use_spatial_splits = true;
    static constexpr int kLeafMax = 8;

    int build_tree() {
        const int n = int(v0.size());
        std::vector<Ref> refs(n);
        Box all;
        for (int i = 0; i < n; i++) {
            refs[i].id = i;
            refs[i].box.grow(v0[i]); refs[i].box.grow(v1[i]); refs[i].box.grow(v2[i]);
            all.grow(refs[i].box);
        }
        root_area = all.half_area();
        order.clear();
        if (const char* e = std::getenv("RODENT_B200_NO_SPATIAL_SPLITS")) use_spatial_splits = e[0] == '0' || e[0] == 0;
        return build2(refs, all, 0);
    }

    // leaf writer of converter.cpp:207-259
    int write_leaf(const Bvh2Node& leaf) {
        const int first_tri4 = int(tris.size());
        for (int i = 0; i < leaf.count; i += 4) {
            Tri4 t; std::memset(&t, 0, sizeof t);
            const int c = std::min(4, leaf.count - i);
            for (int j = 0; j < c; j++) {
                const int id = order[leaf.first + i + j];
                const F3 e1 = v0[id] - v1[id], e2 = v2[id] - v0[id], n = cross(e1, e2);
                t.v0[0][j] = v0[id].x; t.v0[1][j] = v0[id].y; t.v0[2][j] = v0[id].z;
                t.e1[0][j] = e1.x; t.e1[1][j] = e1.y; t.e1[2][j] = e1.z;
                t.e2[0][j] = e2.x; t.e2[1][j] = e2.y; t.e2[2][j] = e2.z;
                t.n[0][j] = n.x; t.n[1][j] = n.y; t.n[2][j] = n.z;
                t.prim_id[j] = id; t.geom_id[j] = geom[id];
            }
            for (int j = c; j < 4; j++) t.prim_id[j] = -1;
            tris.push_back(t);
        }
        tris.back().prim_id[3] |= int32_t(0x80000000u);
        return ~first_tri4;
    }

    // node writer of converter.cpp:160-204: collapse the binary tree to the node's arity by always opening
    // the child with the largest area
    int write_node(int root2) {
        std::vector<int> kids{n2[root2].left, n2[root2].right};
        while (int(kids.size()) < kArity) {
            int pick = -1; float area = -1;
            for (size_t i = 0; i < kids.size(); i++)
                if (n2[kids[i]].left >= 0 && n2[kids[i]].box.half_area() > area) { area = n2[kids[i]].box.half_area(); pick = int(i); }
            if (pick < 0) break;
            const int k = kids[pick];
            kids[pick] = n2[k].left;
            kids.push_back(n2[k].right);
        }
        const int id = int(nodes.size());
        nodes.emplace_back();
        std::memset(&nodes[id], 0, sizeof(NodeT));
        for (int j = 0; j < kArity; j++) {
            const float inf = std::numeric_limits<float>::infinity();
            if (j < int(kids.size())) {
                const Box& b = n2[kids[j]].box;
                nodes[id].bounds[0][j] = b.lo.x; nodes[id].bounds[1][j] = b.hi.x;
                nodes[id].bounds[2][j] = b.lo.y; nodes[id].bounds[3][j] = b.hi.y;
                nodes[id].bounds[4][j] = b.lo.z; nodes[id].bounds[5][j] = b.hi.z;
            } else {
                for (int r = 0; r < 6; r += 2) { nodes[id].bounds[r][j] = inf; nodes[id].bounds[r + 1][j] = -inf; }
            }
        }
        for (size_t j = 0; j < kids.size(); j++) {
            const int child = n2[kids[j]].left >= 0 ? write_node(kids[j]) + 1 : write_leaf(n2[kids[j]]);
            nodes[id].child[j] = child;
        }
        return id;
    }

    // The binary tree itself as Node2 / Tri1 (src/traversal/mapping_gpu.impala:3-16): one Tri1 per triangle, the last of
    // a leaf flagged in prim_id's sign bit; node ids are 1-based, the root is node 1 and always an inner node.
    int emit2_leaf(const Bvh2Node& leaf, std::vector<Tri1>& out) const {
        const int first = int(out.size());
        for (int i = 0; i < leaf.count; i++) {
            const int id = order[leaf.first + i];
            const F3 e1 = v0[id] - v1[id], e2 = v2[id] - v0[id];
            Tri1 t{};
            t.v0[0] = v0[id].x; t.v0[1] = v0[id].y; t.v0[2] = v0[id].z;
            t.e1[0] = e1.x; t.e1[1] = e1.y; t.e1[2] = e1.z; t.e2[0] = e2.x; t.e2[1] = e2.y; t.e2[2] = e2.z;
            t.geom_id = geom[id]; t.prim_id = id;
            out.push_back(t);
        }
        out.back().prim_id |= int32_t(0x80000000u);
        return ~first;
    }
    int emit2_node(int k, std::vector<Node2>& out_nodes, std::vector<Tri1>& out_tris) const {
        const int id = int(out_nodes.size());
        out_nodes.emplace_back();
        const int kids[2] = {n2[k].left, n2[k].right};
        for (int j = 0; j < 2; j++) {
            const Box& b = n2[kids[j]].box;
            float* dst = out_nodes[id].bounds + 6 * j;
            dst[0] = b.lo.x; dst[1] = b.hi.x; dst[2] = b.lo.y; dst[3] = b.hi.y; dst[4] = b.lo.z; dst[5] = b.hi.z;
        }
        for (int j = 0; j < 2; j++) {
            const int c = n2[kids[j]].left >= 0 ? emit2_node(kids[j], out_nodes, out_tris) + 1 : emit2_leaf(n2[kids[j]], out_tris);
            out_nodes[id].child[j] = c;
            out_nodes[id].pad[j] = 0;
        }
        return id;
    }
    void run2(std::vector<Node2>& out_nodes, std::vector<Tri1>& out_tris) {
        const int root = build_tree();
        if (n2[root].left < 0) {
            // a scene of one leaf: the root points at it twice (a triangle found twice is found at the same t)
            Node2 node{};
            for (int j = 0; j < 2; j++) {
                const Box& b = n2[root].box;
                float* dst = node.bounds + 6 * j;
                dst[0] = b.lo.x; dst[1] = b.hi.x; dst[2] = b.lo.y; dst[3] = b.hi.y; dst[4] = b.lo.z; dst[5] = b.hi.z;
            }
            out_nodes.push_back(node);
            const int leaf = emit2_leaf(n2[root], out_tris);
            out_nodes[0].child[0] = out_nodes[0].child[1] = leaf;
        } else {
            emit2_node(root, out_nodes, out_tris);
        }
    }

    void run() {
        const int root = build_tree();
        if (n2[root].left < 0) {
            // a single leaf: the root node (id 1) must still be an inner node
            nodes.emplace_back();
            std::memset(&nodes[0], 0, sizeof(NodeT));
            const float inf = std::numeric_limits<float>::infinity();
            for (int j = 0; j < kArity; j++)
                for (int r = 0; r < 6; r += 2) { nodes[0].bounds[r][j] = inf; nodes[0].bounds[r + 1][j] = -inf; }
            const Box& b = n2[root].box;
            nodes[0].bounds[0][0] = b.lo.x; nodes[0].bounds[1][0] = b.hi.x; nodes[0].bounds[2][0] = b.lo.y;
            nodes[0].bounds[3][0] = b.hi.y; nodes[0].bounds[4][0] = b.lo.z; nodes[0].bounds[5][0] = b.hi.z;
            nodes[0].child[0] = write_leaf(n2[root]);
        } else {
            write_node(root);
        }
    }
};

void put4(std::vector<float>& dst, F3 v) { dst.push_back(v.x); dst.push_back(v.y); dst.push_back(v.z); dst.push_back(0.0f); }

// lights + light ids from emissive materials: converter.cpp:770-818, 832
void collect_lights(Scene& s, const std::vector<F3>& verts) {
    const int num_tris = int(s.indices.size() / 4);
    s.light_ids.assign(num_tris, 0);
    for (int i = 0; i < num_tris; i++) {
        const int m = s.indices[4 * i + 3];
        if (!s.materials[m].is_emissive) continue;
        const F3 a = verts[s.indices[4 * i]], b = verts[s.indices[4 * i + 1]], c = verts[s.indices[4 * i + 2]];
        F3 n = cross(b - a, c - a);
        // a triangle without area (or one no Tri4 ever described: its vertices are zero) cannot be sampled: it would
        // carry inv_area = inf and a NaN normal into every next-event estimate that picks it
        if (!(length(n) > 0.0f)) continue;
        RodentLight l{};
        l.inv_area = 1.0f / (0.5f * length(n));
        n = normalize(n);
        l.v0[0] = a.x; l.v0[1] = a.y; l.v0[2] = a.z; l.v1[0] = b.x; l.v1[1] = b.y; l.v1[2] = b.z; l.v2[0] = c.x; l.v2[1] = c.y; l.v2[2] = c.z;
        l.n[0] = n.x; l.n[1] = n.y; l.n[2] = n.z;
        for (int k = 0; k < 3; k++) l.color[k] = s.materials[m].ke[k];
        l.prim = i; l.map_ke = s.materials[m].map_ke;
        s.light_ids[i] = int(s.lights.size());
        s.lights.push_back(l);
    }
}

}  // namespace

constexpr int kBvh2MinTriangles = 4096;

// OBJ + MTL -> the scene's material table and images (the first half of convert_obj, converter.cpp:565-602, 748-768, 870-913);
// `obj` comes back with its faces' material ids rewritten to the cleaned-up table.
static bool load_materials(const std::string& path, ObjFile& obj, Scene* scene) {
    if (!parse_obj(path, obj)) return false;
    std::map<std::string, ObjMaterial> lib;
    const size_t slash = path.find_last_of('/');
    const std::string dir = slash == std::string::npos ? "." : path.substr(0, slash);
    for (auto& name : obj.mtl_libs)
        if (!parse_mtl(dir + "/" + name, lib)) return false;
    cleanup_materials(obj, lib);

    // Images, converter.cpp:595-602, 748-768: one per distinct file name, relative to the OBJ's directory, '\\' -> '/';
    // .png, .jpg / .jpeg and .tga are decoded, an unknown extension gives the reference's 1x1 black dummy image.  (The reference registers map_Ks
    // under map_Kd's name, :600, so a specular map silently samples image 0 there; here it samples its own file.)
    std::unordered_map<std::string, int> images;
    bool failed = false;
    auto image_of = [&](std::string name, const std::string& material) -> int {
        if (name.empty()) return 0;
        std::replace(name.begin(), name.end(), '\\', '/');
        auto it = images.find(name);
        if (it != images.end()) return it->second;
        auto ends_with = [&](const char* ext) { const size_t n = std::strlen(ext); return name.size() >= n && name.compare(name.size() - n, n, ext) == 0; };
        int id = 0;
        if (ends_with(".png")) {
            int w = 0, h = 0; std::vector<uint32_t> px; std::string why;
            if (load_png(dir + "/" + name, w, h, px, why)) id = scene->add_texture(px.data(), w, h);
            else { fail("cannot load PNG file '" + dir + "/" + name + "': " + why); failed = true; }
        } else if (ends_with(".jpg") || ends_with(".jpeg")) {
            int w = 0, h = 0; std::vector<uint32_t> px; std::string why;
            if (load_jpg(dir + "/" + name, w, h, px, why)) id = scene->add_texture(px.data(), w, h);
            else if (why == "cannot open file") { fail("cannot load JPG file '" + dir + "/" + name + "': " + why); failed = true; }
            else warn("'" + name + "' (material '" + material + "'): " + why + "; the material's constant colour is used instead");
        } else if (ends_with(".tga")) {
            int w = 0, h = 0; std::vector<uint32_t> px; std::string why;
            if (load_tga(dir + "/" + name, w, h, px, why)) id = scene->add_texture(px.data(), w, h);
            else if (why == "cannot open file") { fail("cannot load TGA file '" + dir + "/" + name + "': " + why); failed = true; }
            else warn("'" + name + "' (material '" + material + "'): " + why + "; the material's constant colour is used instead");
        } else if (ends_with(".tiff")) {
            warn("no decoder for '" + name + "' (material '" + material + "'): the material's constant colour is used instead");
        } else {
            const uint32_t black = 0xFF000000u;
            id = scene->add_texture(&black, 1, 1);
        }
        return images[name] = id;
    };
    for (auto& name : obj.materials) {
        const ObjMaterial& m = lib[name];
        RodentMaterial r = make_material(m);
        r.map_ke = image_of(m.map_ke, name);            // converter.cpp:794-801: the light's colour is the texture
        if (r.bsdf == RODENT_BSDF_DIFFUSE || r.bsdf == RODENT_BSDF_PHONG || r.bsdf == RODENT_BSDF_MIX) {
            r.map_kd = image_of(m.map_kd, name);
            r.map_ks = image_of(m.map_ks, name);
        }
        scene->materials.push_back(r);
        scene->material_names.push_back(name);
    }
    return !failed;
}

Scene* load_obj_scene(const std::string& path) {
    ObjFile obj;
    auto scene = new Scene();
    if (!load_materials(path, obj, scene)) { delete scene; return nullptr; }

    // compute_tri_mesh, obj.cpp:412-509: per object, vertices de-duplicated by (v, t, n) in order of first use
    std::vector<F3> verts, normals, face_normals;
    std::vector<std::array<float, 2>> uvs;
    for (auto& faces : obj.objects) {
        std::map<std::array<int, 3>, int> mapping;
        std::vector<std::array<int, 3>> keys;
        bool has_normals = false, has_uvs = false;
        auto index_of = [&](const ObjIndex& i) {
            const std::array<int, 3> k{i.v, i.t, i.n};
            auto it = mapping.find(k);
            if (it != mapping.end()) return it->second;
            has_normals |= i.n != 0; has_uvs |= i.t != 0;
            const int id = int(keys.size());
            mapping.emplace(k, id); keys.push_back(k);
            return id;
        };
        const size_t vtx_offset = verts.size(), tri_offset = scene->indices.size() / 4;
        std::vector<std::array<int, 4>> tri_list;
        for (auto& f : faces) {
            std::vector<int> ids;
            for (auto& i : f.idx) ids.push_back(index_of(i));
            for (size_t i = 1; i + 1 < ids.size(); i++) tri_list.push_back({ids[0], ids[i], ids[i + 1], f.material});   // fan
        }
        if (tri_list.empty()) continue;
        for (auto& k : keys) {
            verts.push_back(obj.vertices[k[0]]);
            uvs.push_back(has_uvs ? obj.texcoords[k[1]] : std::array<float, 2>{0, 0});
            normals.push_back(has_normals ? obj.normals[k[2]] : F3{});
        }
        for (auto& t : tri_list) {
            for (int c = 0; c < 3; c++) scene->indices.push_back(int(t[c] + vtx_offset));
            scene->indices.push_back(t[3]);
            const F3 a = verts[t[0] + vtx_offset], b = verts[t[1] + vtx_offset], c = verts[t[2] + vtx_offset];
            face_normals.push_back(normalize(cross(b - a, c - a)));            // obj.cpp:385-395
        }
        if (!has_normals)                                                        // obj.cpp:397-410, 487-492
            for (size_t i = 0; i < tri_list.size(); i++)
                for (int c = 0; c < 3; c++) normals[tri_list[i][c] + vtx_offset] = normals[tri_list[i][c] + vtx_offset] + face_normals[tri_offset + i];
    }
    for (auto& n : normals) {                                                    // obj.cpp:495-504
        const float len2 = dot(n, n);
        if (len2 <= std::numeric_limits<float>::epsilon() || std::isnan(len2)) n = {0.0f, 1.0f, 0.0f};
        else n = n * (1.0f / std::sqrt(len2));
    }
    if (scene->indices.empty()) { delete scene; fail("'" + path + "' contains no triangle"); return nullptr; }

    for (auto& v : verts) put4(scene->vertices, v);
    for (auto& n : normals) put4(scene->normals, n);
    for (auto& n : face_normals) put4(scene->face_normals, n);
    for (auto& t : uvs) { scene->texcoords.push_back(t[0]); scene->texcoords.push_back(t[1]); scene->texcoords.push_back(0); scene->texcoords.push_back(0); }
    collect_lights(*scene, verts);

    const int num_tris = int(scene->indices.size() / 4);
    std::vector<F3> a(num_tris), b(num_tris), c(num_tris); std::vector<int> geom(num_tris);
    for (int i = 0; i < num_tris; i++) {
        a[i] = verts[scene->indices[4 * i]]; b[i] = verts[scene->indices[4 * i + 1]]; c[i] = verts[scene->indices[4 * i + 2]];
        geom[i] = scene->indices[4 * i + 3];
    }
    Builder<Node8> builder{a, b, c, geom, {}, {}, {}, {}, scene->nodes, scene->tris};
    builder.run();
    // Scenes of some size also get the BVH2 / Tri1 the reference's GPU device renders from (mapping_gpu.impala:505-509): the
    // render loop is 1.7x faster through it on Sponza; a handful of triangles (the Cornell box) gains nothing.
    if (num_tris >= kBvh2MinTriangles) build_bvh2(*scene);
    return scene;
}

// The same triangles under a BVH4 (for the .bvh writer: the reference's files carry a BVH4 and a BVH8 block,
// tools/bvh_extractor/extract_bvh4_8.cpp:9-42).
void build_bvh4(Scene& scene) {
    if (!scene.nodes4.empty()) return;
    const int num_tris = int(scene.indices.size() / 4);
    std::vector<F3> a(num_tris), b(num_tris), c(num_tris); std::vector<int> geom(num_tris);
    for (int i = 0; i < num_tris; i++) {
        const int* idx = &scene.indices[4 * i];
        auto vert = [&](int k) { return F3{scene.vertices[4 * k], scene.vertices[4 * k + 1], scene.vertices[4 * k + 2]}; };
        a[i] = vert(idx[0]); b[i] = vert(idx[1]); c[i] = vert(idx[2]); geom[i] = idx[3];
    }
    Builder<Node4> builder{a, b, c, geom, {}, {}, {}, {}, scene.nodes4, scene.tris4};
    builder.run();
}

// Replaces the scene's BVH8 by one from this repository's builder over the scene's own triangles (a scene made with
// scene_from_bvh8 carries the caller's tree until then).
void rebuild_bvh8(Scene& scene) {
    const int num_tris = int(scene.indices.size() / 4);
    std::vector<F3> a(num_tris), b(num_tris), c(num_tris); std::vector<int> geom(num_tris);
    for (int i = 0; i < num_tris; i++) {
        const int* idx = &scene.indices[4 * i];
        auto vert = [&](int k) { return F3{scene.vertices[4 * k], scene.vertices[4 * k + 1], scene.vertices[4 * k + 2]}; };
        a[i] = vert(idx[0]); b[i] = vert(idx[1]); c[i] = vert(idx[2]); geom[i] = idx[3];
    }
    scene.nodes.clear(); scene.tris.clear();
    Builder<Node8> builder{a, b, c, geom, {}, {}, {}, {}, scene.nodes, scene.tris};
    builder.run();
}

void build_bvh2(Scene& scene) {
    if (!scene.nodes2.empty()) return;
    const int num_tris = int(scene.indices.size() / 4);
    std::vector<F3> a(num_tris), b(num_tris), c(num_tris); std::vector<int> geom(num_tris);
    for (int i = 0; i < num_tris; i++) {
        const int* idx = &scene.indices[4 * i];
        auto vert = [&](int k) { return F3{scene.vertices[4 * k], scene.vertices[4 * k + 1], scene.vertices[4 * k + 2]}; };
        a[i] = vert(idx[0]); b[i] = vert(idx[1]); c[i] = vert(idx[2]); geom[i] = idx[3];
    }
    std::vector<Node8> unused_nodes; std::vector<Tri4> unused_tris;
    Builder<Node8> builder{a, b, c, geom, {}, {}, {}, {}, unused_nodes, unused_tris};
    builder.run2(scene.nodes2, scene.tris1);
}

bool set_bvh2(Scene& scene, const Node2* nodes, int num_nodes, const Tri1* tris, int num_tri1) {
    const int num_prims = int(scene.indices.size() / 4);
    if (num_nodes <= 0 || num_tri1 <= 0) return fail("empty BVH2");
    std::vector<Tri1> t(tris, tris + num_tri1);
    for (auto& tri : t) {
        const int p = tri.prim_id & 0x7FFFFFFF;
        if (p >= num_prims) return fail("prim_id out of range in BVH2");
        tri.geom_id = scene.indices[4 * p + 3];
        if (tri.geom_id < 0 || tri.geom_id >= int(scene.materials.size())) return fail("geometry id out of range in BVH2");
    }
    scene.nodes2.assign(nodes, nodes + num_nodes);
    scene.tris1 = std::move(t);
    return true;
}

Scene* scene_from_bvh8(const Node8* nodes, int num_nodes, const Tri4* tris, int num_tri4,
                       const RodentMaterial* materials, int num_materials, const int32_t* material_of_prim, int num_prims) {
    if (num_nodes <= 0 || num_tri4 < 0 || num_materials <= 0 || num_prims < 0 || !nodes || !tris || !materials || !material_of_prim) {
        fail("scene_from_bvh8: empty or null input");
        return nullptr;
    }
    // the ids index the materials table, the per-material bins of the sort and the light table on the device
    for (int p = 0; p < num_prims; p++)
        if (material_of_prim[p] < 0 || material_of_prim[p] >= num_materials) {
            std::fprintf(stderr, "rodent_b200: material_of_prim[%d] = %d, the scene has %d materials\n", p, material_of_prim[p], num_materials);
            return nullptr;
        }
    auto scene = new Scene();
    scene->nodes.assign(nodes, nodes + num_nodes);
    scene->tris.assign(tris, tris + num_tri4);
    scene->materials.assign(materials, materials + num_materials);
    scene->material_names.resize(num_materials);
    std::vector<F3> verts(size_t(num_prims) * 3), face_normals(num_prims);
    std::vector<char> seen(num_prims, 0);
    scene->indices.assign(size_t(num_prims) * 4, 0);
    for (int k = 0; k < num_tri4; k++)
        for (int j = 0; j < 4; j++) {
            Tri4& t = scene->tris[k];
            if (t.prim_id[j] == -1) continue;
            const int p = t.prim_id[j] & 0x7FFFFFFF;
            if (p >= num_prims) { delete scene; fail("prim_id out of range in BVH"); return nullptr; }
            t.geom_id[j] = material_of_prim[p];
            if (seen[p]) continue;
            seen[p] = 1;
            const F3 v0{t.v0[0][j], t.v0[1][j], t.v0[2][j]}, e1{t.e1[0][j], t.e1[1][j], t.e1[2][j]}, e2{t.e2[0][j], t.e2[1][j], t.e2[2][j]};
            const F3 v1 = v0 - e1, v2 = v0 + e2;                                  // make_tri, intersection.impala:110-119
            verts[3 * p] = v0; verts[3 * p + 1] = v1; verts[3 * p + 2] = v2;
            const F3 n = cross(v1 - v0, v2 - v0);
            const float len = length(n);
            face_normals[p] = len > 0 ? n * (1.0f / len) : F3{0, 1, 0};
            scene->indices[4 * p] = 3 * p; scene->indices[4 * p + 1] = 3 * p + 1; scene->indices[4 * p + 2] = 3 * p + 2;
            scene->indices[4 * p + 3] = material_of_prim[p];
        }
    for (int p = 0; p < num_prims; p++) {
        for (int c = 0; c < 3; c++) { put4(scene->vertices, verts[3 * p + c]); put4(scene->normals, face_normals[p]); }
        put4(scene->face_normals, face_normals[p]);
        if (!seen[p]) scene->indices[4 * p + 3] = material_of_prim[p];
    }
    scene->texcoords.assign(size_t(num_prims) * 12, 0.0f);
    collect_lights(*scene, verts);
    return scene;
}

// ---- the converter's data/ directory (convert_obj, src/driver/converter.cpp:403-438, 682-745) ---------------------------
// What the reference's converter leaves next to its generated main.impala: the mesh as LZ4-framed buffers
// (data/vertices.bin, normals.bin, face_normals.bin, texcoords.bin -- elements padded to 16 bytes for the GPU targets --
// and indices.bin, four ints per triangle, the fourth the geometry = material id), the BVH of the target's layout in
// data/bvh.bin, and data/bvh.stamp naming the target and the OBJ file.  Materials and lights are NOT in there: the
// converter prints them into main.impala as code.  Here they come from the OBJ / MTL the stamp names (the same
// cleanup rules give the same ids); with --fusion the per-triangle simple_kd / ks / ns buffers are turned back into
// table entries.
extern "C" void* rodent_b200_load_buffer(const char* file, int64_t* size);
extern "C" void rodent_b200_free_buffer(void* p);
extern "C" int32_t rodent_b200_write_buffer(const char* file, const void* data, int64_t bytes);
extern "C" int32_t rodent_b200_load_bvh_bin(const char* file, int32_t node_size, int32_t tri_size, void** nodes, int64_t* num_nodes, void** tris, int64_t* num_tris);
extern "C" int32_t rodent_b200_append_bvh_bin(const char* file, int32_t node_size, int32_t tri_size, const void* nodes, int64_t num_nodes, const void* tris, int64_t num_tris);

namespace {
bool read_buffer_file(const std::string& file, std::vector<uint8_t>& out) {
    int64_t n = 0;
    void* p = rodent_b200_load_buffer(file.c_str(), &n);
    if (!p) return false;
    out.assign(static_cast<uint8_t*>(p), static_cast<uint8_t*>(p) + n);
    rodent_b200_free_buffer(p);
    return true;
}
// elements of `comps` floats, stored with a stride of `comps` or 4 floats -> float4 stride
bool unpack_vectors(const std::vector<uint8_t>& raw, size_t count, int comps, std::vector<float>& out) {
    if (count == 0 || raw.size() % (count * 4) != 0) return false;
    const size_t stride = raw.size() / (count * 4);
    if (stride != size_t(comps) && stride != 4) return false;
    const float* src = reinterpret_cast<const float*>(raw.data());
    out.assign(count * 4, 0.0f);
    for (size_t i = 0; i < count; i++)
        for (int c = 0; c < comps; c++) out[4 * i + c] = src[stride * i + c];
    return true;
}
template <typename NodeT, typename TriT>
bool load_bvh_block(const std::string& file, std::vector<NodeT>& nodes, std::vector<TriT>& tris) {
    void* n = nullptr; void* t = nullptr; int64_t nn = 0, nt = 0;
    std::FILE* probe = std::fopen(file.c_str(), "rb");
    if (!probe) return false;
    std::fclose(probe);
    // (rodent_b200_load_bvh_bin reports a missing layout on stderr; probing three layouts, two of them are expected to miss)
    std::fflush(stderr);
    if (!rodent_b200_load_bvh_bin(file.c_str(), int32_t(sizeof(NodeT)), int32_t(sizeof(TriT)), &n, &nn, &t, &nt)) return false;
    nodes.assign(static_cast<NodeT*>(n), static_cast<NodeT*>(n) + nn);
    tris.assign(static_cast<TriT*>(t), static_cast<TriT*>(t) + nt);
    rodent_b200_free_buffer(n); rodent_b200_free_buffer(t);
    return true;
}
}  // namespace

Scene* load_data_dir(const std::string& dir, const std::string& obj_path_arg) {
    std::string obj_path = obj_path_arg;
    if (obj_path.empty()) {                                   // data/bvh.stamp: "<target> <obj file>", converter.cpp:741-742
        std::ifstream stamp(dir + "/bvh.stamp");
        int target = 0;
        if (!(stamp >> target) || !std::getline(stamp >> std::ws, obj_path) || obj_path.empty()) {
            fail("'" + dir + "/bvh.stamp' does not name the OBJ file; pass it explicitly");
            return nullptr;
        }
    }
    auto scene = new Scene();
    ObjFile obj;
    if (!load_materials(obj_path, obj, scene)) { delete scene; return nullptr; }

    std::vector<uint8_t> raw_idx, raw_v, raw_n, raw_fn, raw_t;
    if (!read_buffer_file(dir + "/indices.bin", raw_idx) || !read_buffer_file(dir + "/vertices.bin", raw_v) ||
        !read_buffer_file(dir + "/normals.bin", raw_n) || !read_buffer_file(dir + "/face_normals.bin", raw_fn) ||
        !read_buffer_file(dir + "/texcoords.bin", raw_t)) { delete scene; return nullptr; }
    if (raw_idx.empty() || raw_idx.size() % 16) { delete scene; fail("indices.bin is not a list of int4"); return nullptr; }
    const int num_tris = int(raw_idx.size() / 16);
    scene->indices.assign(reinterpret_cast<const int32_t*>(raw_idx.data()), reinterpret_cast<const int32_t*>(raw_idx.data()) + size_t(num_tris) * 4);
    int num_verts = 0;
    for (int i = 0; i < num_tris; i++)
        for (int c = 0; c < 3; c++) {
            if (scene->indices[4 * i + c] < 0) { delete scene; fail("negative vertex index in indices.bin"); return nullptr; }
            num_verts = std::max(num_verts, scene->indices[4 * i + c] + 1);
        }
    // the arrays may hold vertices no triangle uses: their length follows from the file sizes.  Padded (GPU targets) or not:
    // a vertex is 16 bytes in both vertices.bin and texcoords.bin when padded, 12 and 8 when not.
    const bool padded = raw_v.size() == raw_t.size();
    const size_t nv_guess = raw_v.size() / (padded ? 16 : 12);
    const size_t nv = nv_guess >= size_t(num_verts) && raw_v.size() % (padded ? 16 : 12) == 0 ? nv_guess : 0;
    if (nv == 0 || !unpack_vectors(raw_v, nv, 3, scene->vertices) || !unpack_vectors(raw_n, nv, 3, scene->normals) ||
        !unpack_vectors(raw_fn, size_t(num_tris), 3, scene->face_normals) || !unpack_vectors(raw_t, nv, 2, scene->texcoords)) {
        delete scene; fail("mesh buffers of '" + dir + "' do not agree in size"); return nullptr;
    }

    // --fusion: every simple material became geometry `num_complex`, its colours moved to per-triangle buffers
    // (converter.cpp:682-709).  One table entry per distinct (kd, ks, ns) brings them back.
    std::vector<uint8_t> raw_kd, raw_ks, raw_ns;
    {
        std::FILE* f = std::fopen((dir + "/simple_kd.bin").c_str(), "rb");
        if (f) {
            std::fclose(f);
            std::vector<float> kd, ks;
            if (!read_buffer_file(dir + "/simple_kd.bin", raw_kd) || !read_buffer_file(dir + "/simple_ks.bin", raw_ks) ||
                !read_buffer_file(dir + "/simple_ns.bin", raw_ns) || !unpack_vectors(raw_kd, size_t(num_tris), 3, kd) ||
                !unpack_vectors(raw_ks, size_t(num_tris), 3, ks) || raw_ns.size() != size_t(num_tris) * 4) {
                delete scene; fail("simple_kd / simple_ks / simple_ns do not match the mesh"); return nullptr;
            }
            int fused = 0;                                     // the fused geometry is the last id in use (= num_complex)
            for (int i = 0; i < num_tris; i++) fused = std::max(fused, scene->indices[4 * i + 3]);
            if (fused > int(scene->materials.size())) { delete scene; fail("indices.bin names more geometries than the OBJ has materials"); return nullptr; }
            const float* ns = reinterpret_cast<const float*>(raw_ns.data());
            std::map<std::array<float, 7>, int> seen;
            scene->materials.resize(size_t(fused));
            scene->material_names.resize(size_t(fused));
            for (int i = 0; i < num_tris; i++) {
                if (scene->indices[4 * i + 3] != fused) continue;
                const std::array<float, 7> key{kd[4 * i], kd[4 * i + 1], kd[4 * i + 2], ks[4 * i], ks[4 * i + 1], ks[4 * i + 2], ns[i]};
                auto it = seen.find(key);
                if (it == seen.end()) {
                    ObjMaterial m;
                    m.kd = {key[0], key[1], key[2]}; m.ks = {key[3], key[4], key[5]}; m.ns = key[6]; m.ni = 1.0f; m.illum = 2;
                    it = seen.emplace(key, int(scene->materials.size())).first;
                    scene->materials.push_back(make_material(m));
                    scene->material_names.push_back("fused");
                }
                scene->indices[4 * i + 3] = it->second;
            }
        }
    }
    const int num_materials = int(scene->materials.size());
    for (int i = 0; i < num_tris; i++)
        if (scene->indices[4 * i + 3] < 0 || scene->indices[4 * i + 3] >= num_materials) {
            delete scene; fail("indices.bin names geometry " + std::to_string(scene->indices[4 * i + 3]) + ", '" + obj_path + "' has " + std::to_string(num_materials) + " materials");
            return nullptr;
        }

    std::vector<F3> verts(nv);
    for (size_t i = 0; i < nv; i++) verts[i] = {scene->vertices[4 * i], scene->vertices[4 * i + 1], scene->vertices[4 * i + 2]};
    collect_lights(*scene, verts);

    // The BVH the converter wrote, whatever its layout; the geometry ids inside it are refreshed from indices.bin (fusion
    // renumbered them).  A BVH8 is adopted as the scene's, a BVH2 / Tri1 as its second tree, and whatever is missing is
    // built from the triangles.
    const std::string bvh_file = dir + "/bvh.bin";
    bool have8 = load_bvh_block(bvh_file, scene->nodes, scene->tris);
    if (have8) {
        for (Tri4& t : scene->tris)
            for (int j = 0; j < 4; j++) {
                if (t.prim_id[j] == -1) continue;
                const int p = t.prim_id[j] & 0x7FFFFFFF;
                if (p >= num_tris) { delete scene; fail("prim_id out of range in bvh.bin"); return nullptr; }
                t.geom_id[j] = scene->indices[4 * p + 3];
            }
    }
    std::vector<Node2> n2; std::vector<Tri1> t1;
    const bool have2 = load_bvh_block(bvh_file, n2, t1);
    load_bvh_block(bvh_file, scene->nodes4, scene->tris4);
    if (!have8) {
        std::vector<F3> a(num_tris), b(num_tris), c(num_tris); std::vector<int> geom(num_tris);
        for (int i = 0; i < num_tris; i++) {
            a[i] = verts[scene->indices[4 * i]]; b[i] = verts[scene->indices[4 * i + 1]]; c[i] = verts[scene->indices[4 * i + 2]];
            geom[i] = scene->indices[4 * i + 3];
        }
        Builder<Node8> builder{a, b, c, geom, {}, {}, {}, {}, scene->nodes, scene->tris};
        builder.run();
    }
    if (have2 && !set_bvh2(*scene, n2.data(), int(n2.size()), t1.data(), int(t1.size()))) { delete scene; return nullptr; }
    if (!have2 && num_tris >= kBvh2MinTriangles) build_bvh2(*scene);
    return scene;
}

// convert_obj's data/ directory from a loaded scene: `arity` 2, 4 or 8 picks the layout written to bvh.bin (the reference
// writes the one of its target: 2 for the GPU targets, 4 for generic / sse42 / asimd, 8 for avx / avx2), `padded` the
// 16-byte elements of the GPU targets (converter.cpp:630-633).
bool write_data_dir(Scene& scene, const std::string& dir, int arity, bool padded, const std::string& obj_path) {
    const size_t nv = scene.vertices.size() / 4, nt = scene.indices.size() / 4;
    auto pack = [&](const std::vector<float>& src, size_t count, int comps) {
        const int stride = padded ? 4 : comps;
        std::vector<float> out(count * size_t(stride), 0.0f);
        for (size_t i = 0; i < count; i++)
            for (int c = 0; c < comps; c++) out[i * stride + c] = src[4 * i + c];
        return out;
    };
    auto put = [&](const char* name, const std::vector<float>& v) {
        return rodent_b200_write_buffer((dir + "/" + name).c_str(), v.data(), int64_t(v.size() * sizeof(float))) != 0;
    };
    if (!put("vertices.bin", pack(scene.vertices, nv, 3)) || !put("normals.bin", pack(scene.normals, nv, 3)) ||
        !put("face_normals.bin", pack(scene.face_normals, nt, 3)) || !put("texcoords.bin", pack(scene.texcoords, nv, 2)) ||
        !rodent_b200_write_buffer((dir + "/indices.bin").c_str(), scene.indices.data(), int64_t(scene.indices.size() * sizeof(int32_t))))
        return fail("cannot write the mesh buffers into '" + dir + "'");
    const std::string bvh_file = dir + "/bvh.bin";
    std::remove(bvh_file.c_str());
    bool ok = false;
    if (arity == 8) ok = rodent_b200_append_bvh_bin(bvh_file.c_str(), sizeof(Node8), sizeof(Tri4), scene.nodes.data(), int64_t(scene.nodes.size()), scene.tris.data(), int64_t(scene.tris.size()));
    else if (arity == 4) { build_bvh4(scene); ok = rodent_b200_append_bvh_bin(bvh_file.c_str(), sizeof(Node4), sizeof(Tri4), scene.nodes4.data(), int64_t(scene.nodes4.size()), scene.tris4.data(), int64_t(scene.tris4.size())); }
    else if (arity == 2) { build_bvh2(scene); ok = rodent_b200_append_bvh_bin(bvh_file.c_str(), sizeof(Node2), sizeof(Tri1), scene.nodes2.data(), int64_t(scene.nodes2.size()), scene.tris1.data(), int64_t(scene.tris1.size())); }
    if (!ok) return fail("cannot write '" + bvh_file + "'");
    std::ofstream stamp(dir + "/bvh.stamp");
    stamp << (arity == 2 ? 6 : arity == 4 ? 0 : 1) << " " << obj_path;          // Target enum of converter.cpp:24-36: generic 0, avx2 1, nvvm-streaming 6
    return bool(stamp);
}

int Scene::add_texture(const uint32_t* rgba, int width, int height) {
    textures.push_back({width, height, int64_t(texture_pixels.size())});
    texture_pixels.insert(texture_pixels.end(), rgba, rgba + size_t(width) * height);
    return int(textures.size());
}

}  // namespace rb200

using rb200::Scene;

extern "C" {

RodentScene* rodent_b200_scene_load_obj(const char* obj_file) {
    return reinterpret_cast<RodentScene*>(rb200::load_obj_scene(obj_file));
}
RodentScene* rodent_b200_scene_load_data(const char* data_dir, const char* obj_file) {
    return reinterpret_cast<RodentScene*>(rb200::load_data_dir(data_dir ? data_dir : "data", obj_file ? obj_file : ""));
}
int32_t rodent_b200_scene_write_data(RodentScene* scene, const char* data_dir, int32_t bvh_arity, int32_t padded, const char* obj_file) {
    if (!scene || !data_dir || (bvh_arity != 2 && bvh_arity != 4 && bvh_arity != 8)) return 0;
    return rb200::write_data_dir(*reinterpret_cast<Scene*>(scene), data_dir, bvh_arity, padded != 0, obj_file ? obj_file : "") ? 1 : 0;
}
RodentScene* rodent_b200_scene_from_bvh8(const Node8* nodes, int32_t num_nodes, const Tri4* tris, int32_t num_tri4,
                                         const RodentMaterial* materials, int32_t num_materials,
                                         const int32_t* material_of_prim, int32_t num_prims) {
    return reinterpret_cast<RodentScene*>(rb200::scene_from_bvh8(nodes, num_nodes, tris, num_tri4, materials, num_materials, material_of_prim, num_prims));
}
void rodent_b200_scene_view(const RodentScene* scene, RodentSceneView* out) {
    const Scene& s = *reinterpret_cast<const Scene*>(scene);
    out->num_tris = int32_t(s.indices.size() / 4); out->num_vertices = int32_t(s.vertices.size() / 4);
    out->num_materials = int32_t(s.materials.size()); out->num_lights = int32_t(s.lights.size());
    out->num_nodes = int32_t(s.nodes.size()); out->num_tri4 = int32_t(s.tris.size());
    out->vertices = s.vertices.data(); out->normals = s.normals.data(); out->face_normals = s.face_normals.data();
    out->texcoords = s.texcoords.data(); out->indices = s.indices.data(); out->light_ids = s.light_ids.data();
    out->materials = s.materials.data(); out->lights = s.lights.data(); out->nodes = s.nodes.data(); out->tris = s.tris.data();
    out->textures = s.textures.data(); out->texture_pixels = s.texture_pixels.data();
    out->num_texture_pixels = int64_t(s.texture_pixels.size()); out->num_textures = int32_t(s.textures.size()); out->pad = 0;
    out->nodes2 = s.nodes2.empty() ? nullptr : s.nodes2.data(); out->tris1 = s.tris1.empty() ? nullptr : s.tris1.data();
    out->num_nodes2 = int32_t(s.nodes2.size()); out->num_tri1 = int32_t(s.tris1.size());
}
void rodent_b200_scene_build_bvh2(RodentScene* scene) { rb200::build_bvh2(*reinterpret_cast<Scene*>(scene)); }
void rodent_b200_scene_rebuild_bvh8(RodentScene* scene) { rb200::rebuild_bvh8(*reinterpret_cast<Scene*>(scene)); }
int32_t rodent_b200_scene_set_bvh2(RodentScene* scene, const Node2* nodes, int32_t num_nodes, const Tri1* tris, int32_t num_tri1) {
    return rb200::set_bvh2(*reinterpret_cast<Scene*>(scene), nodes, num_nodes, tris, num_tri1) ? 1 : 0;
}
int32_t rodent_b200_scene_add_texture(RodentScene* scene, const uint32_t* rgba, int32_t width, int32_t height) {
    if (!scene || !rgba || width <= 0 || height <= 0) return 0;
    return reinterpret_cast<Scene*>(scene)->add_texture(rgba, width, height);
}
int32_t rodent_b200_scene_add_png(RodentScene* scene, const char* png_file) {
    int w = 0, h = 0; std::vector<uint32_t> px; std::string why;
    if (!scene || !rb200::load_png(png_file, w, h, px, why)) {
        std::fprintf(stderr, "rodent_b200: cannot load PNG file '%s': %s\n", png_file, why.c_str());
        return 0;
    }
    return reinterpret_cast<Scene*>(scene)->add_texture(px.data(), w, h);
}
int32_t rodent_b200_scene_add_tga(RodentScene* scene, const char* tga_file) {
    int w = 0, h = 0; std::vector<uint32_t> px; std::string why;
    if (!scene || !rb200::load_tga(tga_file, w, h, px, why)) {
        std::fprintf(stderr, "rodent_b200: cannot load TGA file '%s': %s\n", tga_file, why.c_str());
        return 0;
    }
    return reinterpret_cast<Scene*>(scene)->add_texture(px.data(), w, h);
}
int32_t rodent_b200_scene_add_jpg(RodentScene* scene, const char* jpg_file) {
    int w = 0, h = 0; std::vector<uint32_t> px; std::string why;
    if (!scene || !rb200::load_jpg(jpg_file, w, h, px, why)) {
        std::fprintf(stderr, "rodent_b200: cannot load JPG file '%s': %s\n", jpg_file, why.c_str());
        return 0;
    }
    return reinterpret_cast<Scene*>(scene)->add_texture(px.data(), w, h);
}
void rodent_b200_scene_free(RodentScene* scene) { delete reinterpret_cast<Scene*>(scene); }
void rodent_b200_scene_bvh4(RodentScene* scene, const Node4** nodes, int32_t* num_nodes, const Tri4** tris, int32_t* num_tri4) {
    Scene& s = *reinterpret_cast<Scene*>(scene);
    rb200::build_bvh4(s);
    *nodes = s.nodes4.data(); *num_nodes = int32_t(s.nodes4.size());
    *tris = s.tris4.data(); *num_tri4 = int32_t(s.tris4.size());
}

}  // extern "C"
