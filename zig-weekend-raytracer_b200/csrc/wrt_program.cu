// wrt_program.cu — host code only (compiled by nvcc for the shared struct definitions).
//
// Turns the reference's pointer tree (IEntity, src/entity.zig:17-24, passed as wrt_entity records) into
//   * a linear program in DFS pre-order — exactly the order BVHNodeEntity.hit / EntityCollection.hit visit
//     children (entity.zig:286-303, 342-368) — where every bvh_node carries the pc to jump to when culled;
//   * two box sets per node: the reference's own cached box (for WRT_CULL_REFERENCE) and a conservative box
//     recomputed from the primitives in the node's local space (for WRT_CULL_TIGHT);
//   * primitive ids = first-visit DFS order.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <new>
#include <stdexcept>
#include <system_error>
#include <thread>

#include "wrt_program.h"
#include "wrt_treebuild.cuh"

namespace wrt {
namespace {

struct Box3 {
    double mn[3], mx[3];
    void reset() {
        for (int k = 0; k < 3; ++k) { mn[k] = std::numeric_limits<double>::infinity(); mx[k] = -std::numeric_limits<double>::infinity(); }
    }
    void grow(const double p[3]) {
        for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); }
    }
    void grow(const Box3& b) {
        for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], b.mn[k]); mx[k] = std::max(mx[k], b.mx[k]); }
    }
    bool valid() const { return mn[0] <= mx[0]; }
};

struct Compiler {
    const wrt_scene* sc;
    CompiledScene& out;
    std::string& err;
    int code = WRT_OK;

    std::vector<Box3> tight;          // per entity, local space; valid flag in tight_state
    std::vector<uint8_t> tight_state; // 0 = not computed, 1 = in progress, 2 = done
    std::vector<uint32_t> prim_id;    // per entity (sphere / quad), WRT_NONE until first visit
    std::vector<uint8_t> on_stack;    // cycle guard for emit()

    Compiler(const wrt_scene* s, CompiledScene& o, std::string& e) : sc(s), out(o), err(e) {}

    bool fail(int c, const std::string& msg) {
        if (code == WRT_OK) { code = c; err = msg; }
        return false;
    }

    bool check_entity(uint32_t id, const char* what) {
        if (id >= sc->n_entities) return fail(WRT_E_INVALID, std::string("entity index out of range in ") + what);
        return true;
    }

    // ---- conservative local-space boxes ---------------------------------------------------------------
    bool tight_box(uint32_t id, Box3& b) {
        if (!check_entity(id, "tight_box")) return false;
        if (tight_state[id] == 2) { b = tight[id]; return true; }
        if (tight_state[id] == 1) return fail(WRT_E_INVALID, "entity graph contains a cycle");
        tight_state[id] = 1;
        const wrt_entity& e = sc->entities[id];
        b.reset();
        switch (e.kind) {
            case WRT_ENT_SPHERE: {
                if (e.a >= sc->n_spheres) return fail(WRT_E_INVALID, "sphere index out of range");
                const wrt_sphere& s = sc->spheres[e.a];
                const double r = std::fabs(s.radius);
                double lo[3], hi[3];
                for (int k = 0; k < 3; ++k) { lo[k] = s.center[k] - r; hi[k] = s.center[k] + r; }
                b.grow(lo); b.grow(hi);
                if (s.is_moving) {
                    for (int k = 0; k < 3; ++k) { lo[k] += s.movement[k]; hi[k] += s.movement[k]; }
                    b.grow(lo); b.grow(hi);
                }
                break;
            }
            case WRT_ENT_QUAD: {
                if (e.a >= sc->n_quads) return fail(WRT_E_INVALID, "quad index out of range");
                const wrt_quad& q = sc->quads[e.a];
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 2; ++j) {
                        double p[3];
                        for (int k = 0; k < 3; ++k) p[k] = q.start[k] + i * q.u[k] + j * q.v[k];
                        b.grow(p);
                    }
                break;
            }
            case WRT_ENT_COLLECTION: {
                if ((uint64_t)e.a + e.b > sc->n_children && e.b) return fail(WRT_E_INVALID, "collection child range out of bounds");
                for (uint32_t k = 0; k < e.b; ++k) {
                    Box3 cb;
                    if (!tight_box(sc->children[e.a + k], cb)) return false;
                    if (cb.valid()) b.grow(cb);
                }
                break;
            }
            case WRT_ENT_BVH_NODE: {
                Box3 l, r;
                if (!tight_box(e.a, l) || !tight_box(e.b, r)) return false;
                if (l.valid()) b.grow(l);
                if (r.valid()) b.grow(r);
                break;
            }
            case WRT_ENT_TRANSLATE: {
                Box3 c;
                if (!tight_box(e.a, c)) return false;
                if (c.valid())
                    for (int k = 0; k < 3; ++k) { b.mn[k] = c.mn[k] + e.p[k]; b.mx[k] = c.mx[k] + e.p[k]; }
                break;
            }
            case WRT_ENT_ROTATE_Y: {
                Box3 c;
                if (!tight_box(e.a, c)) return false;
                if (c.valid()) {
                    const double sn = e.p[0], cs = e.p[1];
                    for (int i = 0; i < 2; ++i)
                        for (int j = 0; j < 2; ++j)
                            for (int k = 0; k < 2; ++k) {
                                const double x = i ? c.mx[0] : c.mn[0], y = j ? c.mx[1] : c.mn[1], z = k ? c.mx[2] : c.mn[2];
                                double p[3] = {cs * x + sn * z, y, -sn * x + cs * z};  // objectToWorldSpace, entity.zig:199-205
                                b.grow(p);
                            }
                }
                break;
            }
            default: return fail(WRT_E_INVALID, "unknown entity kind");
        }
        tight[id] = b;
        tight_state[id] = 2;
        return true;
    }

    // binary32 box for the FP32 culler: padded by 4e-6 of the coordinate magnitude (covers |b*inv| * 2^-22 and the
    // binary64 rounding of the primitives' own arithmetic) and rounded outwards.
    static float round_down(double v) {
        float f = (float)v;
        return ((double)f > v) ? std::nextafterf(f, -std::numeric_limits<float>::infinity()) : f;
    }
    static float round_up(double v) {
        float f = (float)v;
        return ((double)f < v) ? std::nextafterf(f, std::numeric_limits<float>::infinity()) : f;
    }
    static BoxTight padded(const Box3& b) {
        BoxTight t;
        t._p0 = t._p1 = 0.0f;
        if (!b.valid()) {  // empty subtree: a box nothing can hit
            t.min_x = t.min_y = t.min_z = 1.0f;
            t.max_x = t.max_y = t.max_z = -1.0f;
            return t;
        }
        double lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            const double mag = std::max(1.0, std::max(std::fabs(b.mn[k]), std::fabs(b.mx[k])));
            lo[k] = b.mn[k] - mag * 4e-6;
            hi[k] = b.mx[k] + mag * 4e-6;
        }
        t.min_x = round_down(lo[0]); t.min_y = round_down(lo[1]); t.min_z = round_down(lo[2]);
        t.max_x = round_up(hi[0]); t.max_y = round_up(hi[1]); t.max_z = round_up(hi[2]);
        return t;
    }

    uint32_t push_box(const wrt_entity& e, const Box3& tb) {
        BoxRef r;
        r.min_x = e.bbox_min[0]; r.min_y = e.bbox_min[1];
        r.max_x = e.bbox_max[0]; r.max_y = e.bbox_max[1];
        out.boxes_ref.push_back(r);
        out.boxes_tight.push_back(padded(tb));
        Node2 none;
        std::memset(&none, 0, sizeof none);
        none.l_desc = none.r_desc = WRT_NONE;
        out.nodes2.push_back(none);  // filled in for bvh_node ops, stays empty for instance bounds
        return (uint32_t)(out.boxes_ref.size() - 1);
    }

    // ---- program emission ------------------------------------------------------------------------------
    uint32_t nest = 0;  // current nesting of bvh_node / instance ops (bounds the ordered traversal's stack)
    const bool keep_ref_records = keep_reference_trees();
    std::vector<TreeItem>* cur_items = nullptr;  // != null while emitting below a bvh_node; out.tree_inputs collects every reference BVH
    bool emit(uint32_t id, uint32_t xf, uint32_t xf_depth) {
        if (!check_entity(id, "emit")) return false;
        struct Nest { uint32_t& n; uint32_t& mx; bool on; Nest(uint32_t& n_, uint32_t& mx_, bool on_) : n(n_), mx(mx_), on(on_) { if (on) { ++n; if (n > mx) mx = n; } } ~Nest() { if (on) --n; } };
        const uint32_t kind_for_nest = sc->entities[id].kind;
        Nest nest_guard(nest, out.max_nesting, kind_for_nest == WRT_ENT_BVH_NODE || kind_for_nest == WRT_ENT_TRANSLATE || kind_for_nest == WRT_ENT_ROTATE_Y);
        if (on_stack[id]) return fail(WRT_E_INVALID, "entity graph contains a cycle");
        on_stack[id] = 1;
        const wrt_entity& e = sc->entities[id];
        bool ok = true;
        switch (e.kind) {
            case WRT_ENT_SPHERE:
            case WRT_ENT_QUAD: {
                const bool sphere = e.kind == WRT_ENT_SPHERE;
                if (e.a >= (sphere ? sc->n_spheres : sc->n_quads)) { ok = fail(WRT_E_INVALID, "primitive index out of range"); break; }
                const uint32_t mat = sphere ? sc->spheres[e.a].material : sc->quads[e.a].material;
                if (mat >= sc->n_materials) { ok = fail(WRT_E_INVALID, "material index out of range"); break; }
                if (prim_id[id] == WRT_NONE) prim_id[id] = out.n_prims++;
                {  // the op that tests this primitive record (compact stack entries look it up on a hit); one op only, or no compact form
                    std::vector<uint32_t>& table = sphere ? out.sphere_pc : out.quad_pc;
                    if (table[e.a] == WRT_NONE) table[e.a] = (uint32_t)out.ops.size();
                    else out.prim_pc_unique = false;
                }
                out.ops.push_back(make_uint4(sphere ? OP_SPHERE : OP_QUAD, e.a, mat, prim_id[id]));
                break;
            }
            case WRT_ENT_COLLECTION: {
                if (e.c != WRT_NONE) { ok = emit(e.c, xf, xf_depth); break; }  // prefer the BVH (entity.zig:347-349)
                if ((uint64_t)e.a + e.b > sc->n_children && e.b) { ok = fail(WRT_E_INVALID, "collection child range out of bounds"); break; }
                for (uint32_t k = 0; k < e.b && ok; ++k) ok = emit(sc->children[e.a + k], xf, xf_depth);
                break;
            }
            case WRT_ENT_BVH_NODE: {
                Box3 tb;
                if (!tight_box(id, tb)) { ok = false; break; }
                // Does the reference's cached box contain its subtree on the two axes its AABB.hit tests (aabb.zig:80-101)?
                // AABB.offset (aabb.zig:52-60) and RotateY's box (entity.zig:139,143) can make it smaller: the reference
                // then drops hits a conservative culler keeps, and WRT_CULL_AUTO must use the reference's test.
                if (tb.valid() && !(e.bbox_min[0] <= tb.mn[0] && e.bbox_min[1] <= tb.mn[1] && e.bbox_max[0] >= tb.mx[0] &&
                                    e.bbox_max[1] >= tb.mx[1]))
                    ++out.ref_boxes_loose;
                const uint32_t box = push_box(e, tb);
                const size_t at = out.ops.size();
                out.ops.push_back(make_uint4(OP_NODE, box, 0, 0));
                // the leaf entities below this BVH's root are collected for the tree build (out.tree_inputs)
                const bool is_root = (cur_items == nullptr);
                std::vector<TreeItem> root_items;
                if (is_root) cur_items = &root_items;
                auto emit_child = [&](uint32_t child) -> bool {
                    if (!check_entity(child, "bvh_node child")) return false;
                    if (sc->entities[child].kind == WRT_ENT_BVH_NODE) return emit(child, xf, xf_depth);
                    std::vector<TreeItem>* const items = cur_items;
                    cur_items = nullptr;  // a BVH nested inside this leaf (instance, collection) is a tree of its own
                    TreeItem it;
                    Box3 ib;
                    it.start = (uint32_t)out.ops.size();
                    const bool child_ok = emit(child, xf, xf_depth) && tight_box(child, ib);
                    it.end = (uint32_t)out.ops.size();
                    for (int k = 0; k < 3; ++k) { it.mn[k] = ib.mn[k]; it.mx[k] = ib.mx[k]; }
                    cur_items = items;
                    if (child_ok) items->push_back(it);
                    return child_ok;
                };
                const uint32_t l_start = (uint32_t)out.ops.size();
                ok = emit_child(e.a);
                const uint32_t r_start = (uint32_t)out.ops.size();
                // span == 1 nodes hold the same child twice (entity.zig:231-233); the second visit cannot change the result
                if (ok && e.b != e.a) ok = emit_child(e.b);
                const uint32_t end = (uint32_t)out.ops.size();
                out.ops[at].z = end;
                if (is_root) {
                    cur_items = nullptr;
                    if (ok) out.tree_inputs.push_back(TreeInput{box, nest, std::move(root_items)});
                }
                // child-pair record of the reference topology: both children's boxes + where they live in the program.  Only the
                // records the ordered traversal can reach are formed: every root's (a tree of fewer than two leaves is not
                // rebuilt), and all of them when WRT_REFERENCE_TREE=1 keeps the reference topology.
                if (ok && (is_root || keep_ref_records)) {
                    Node2 n2;
                    Box3 lb, rb;
                    lb.reset(); rb.reset();
                    if (!tight_box(e.a, lb)) { ok = false; break; }
                    if (e.b != e.a && !tight_box(e.b, rb)) { ok = false; break; }
                    const BoxTight pl = padded(lb), pr = padded(rb);
                    auto desc = [&](uint32_t start) {
                        return out.ops[start].x == OP_NODE ? (0x80000000u | out.ops[start].y) : start;
                    };
                    n2.lmin[0] = pl.min_x; n2.lmin[1] = pl.min_y; n2.lmin[2] = pl.min_z; n2.l_desc = desc(l_start);
                    n2.lmax[0] = pl.max_x; n2.lmax[1] = pl.max_y; n2.lmax[2] = pl.max_z; n2.l_end = r_start;
                    n2.rmin[0] = pr.min_x; n2.rmin[1] = pr.min_y; n2.rmin[2] = pr.min_z;
                    n2.rmax[0] = pr.max_x; n2.rmax[1] = pr.max_y; n2.rmax[2] = pr.max_z;
                    n2.r_desc = (r_start < end) ? desc(r_start) : WRT_NONE;
                    n2.r_end = end;
                    out.nodes2[box] = n2;
                }
                break;
            }
            case WRT_ENT_TRANSLATE:
            case WRT_ENT_ROTATE_Y: {
                if (xf_depth + 1 > WRT_MAX_XFORM_DEPTH) { ok = fail(WRT_E_LIMIT, "transform nesting deeper than WRT_MAX_XFORM_DEPTH"); break; }
                Box3 tb;
                if (!tight_box(id, tb)) { ok = false; break; }
                // compiler-inserted bound of the instance in its parent's space (ignored by reference culling)
                const uint32_t box = push_box(e, tb);
                const size_t at = out.ops.size();
                out.ops.push_back(make_uint4(OP_NODE_TIGHT_ONLY, box, 0, 0));
                Xform X;
                X.parent = xf;
                if (e.kind == WRT_ENT_TRANSLATE) { X.kind = OP_PUSH_TRANSLATE; X.a = e.p[0]; X.b = e.p[1]; X.c = e.p[2]; }
                else { X.kind = OP_PUSH_ROTATE_Y; X.a = e.p[0]; X.b = e.p[1]; X.c = 0.0; }
                out.xforms.push_back(X);
                const uint32_t me = (uint32_t)(out.xforms.size() - 1);
                out.max_xform_depth = std::max(out.max_xform_depth, xf_depth + 1);
                out.ops.push_back(make_uint4(X.kind, me, 0, 0));
                ok = emit(e.a, me, xf_depth + 1);
                out.ops.push_back(make_uint4(OP_POP, xf, 0, 0));
                out.ops[at].z = (uint32_t)out.ops.size();
                break;
            }
            default: ok = fail(WRT_E_INVALID, "unknown entity kind");
        }
        on_stack[id] = 0;
        return ok;
    }

    bool compile_materials() {
        out.textures.resize(sc->n_textures);
        for (uint32_t i = 0; i < sc->n_textures; ++i) {
            const wrt_texture& t = sc->textures[i];
            Texture T;
            T.r = t.color[0]; T.g = t.color[1]; T.b = t.color[2]; T.inv_scale = t.inv_scale;
            T.kind = t.kind; T.even = t.even; T.odd = t.odd; T.image = t.image;
            if (t.kind == WRT_TEX_CHECKER && (t.even >= sc->n_textures || t.odd >= sc->n_textures))
                return fail(WRT_E_INVALID, "checker texture child out of range");
            if (t.kind == WRT_TEX_IMAGE && t.image >= sc->n_images) return fail(WRT_E_INVALID, "image index out of range");
            if (t.kind > WRT_TEX_IMAGE) return fail(WRT_E_INVALID, "unknown texture kind");
            out.textures[i] = T;
        }
        // checker nesting must terminate within the device loop bound
        for (uint32_t i = 0; i < sc->n_textures; ++i) {
            std::vector<uint32_t> stack{i};
            std::vector<uint32_t> depth{0};
            while (!stack.empty()) {
                uint32_t t = stack.back(), d = depth.back();
                stack.pop_back(); depth.pop_back();
                if (d > 15) return fail(WRT_E_LIMIT, "checker textures nested deeper than 15 (or cyclic)");
                if (sc->textures[t].kind == WRT_TEX_CHECKER) {
                    stack.push_back(sc->textures[t].even); depth.push_back(d + 1);
                    stack.push_back(sc->textures[t].odd); depth.push_back(d + 1);
                }
            }
        }
        out.materials.resize(sc->n_materials);
        for (uint32_t i = 0; i < sc->n_materials; ++i) {
            const wrt_material& m = sc->materials[i];
            Material M;
            M.ar = m.albedo[0]; M.ag = m.albedo[1]; M.ab = m.albedo[2]; M.param = m.param;
            M.kind = m.kind; M.texture = m.texture; M._p0 = M._p1 = 0;
            if (m.kind > WRT_MAT_DIFFUSE_EMISSIVE) return fail(WRT_E_INVALID, "unknown material kind");
            const bool textured = m.kind == WRT_MAT_LAMBERTIAN || m.kind == WRT_MAT_ISOTROPIC || m.kind == WRT_MAT_DIFFUSE_EMISSIVE;
            if (textured && m.texture >= sc->n_textures) return fail(WRT_E_INVALID, "material texture index out of range");
            out.materials[i] = M;
        }
        return true;
    }

    bool compile_geometry() {
        out.spheres.resize(sc->n_spheres);
        for (uint32_t i = 0; i < sc->n_spheres; ++i) {
            const wrt_sphere& s = sc->spheres[i];
            SphereGeom g;
            g.cx = s.center[0]; g.cy = s.center[1]; g.cz = s.center[2]; g.radius = s.radius;
            out.spheres[i] = g;
            if (s.is_moving) out.has_moving = true;
        }
        if (out.has_moving) {
            out.sphere_aux.resize(sc->n_spheres);
            for (uint32_t i = 0; i < sc->n_spheres; ++i) {
                const wrt_sphere& s = sc->spheres[i];
                SphereAux a;
                a.mx = s.movement[0]; a.my = s.movement[1]; a.mz = s.movement[2];
                a.is_moving = s.is_moving; a._pad = 0;
                out.sphere_aux[i] = a;
            }
        }
        out.quads.resize(sc->n_quads);
        for (uint32_t i = 0; i < sc->n_quads; ++i) {
            const wrt_quad& q = sc->quads[i];
            QuadGeom g;
            std::memset(&g, 0, sizeof g);
            g.nx = q.normal[0]; g.ny = q.normal[1]; g.nz = q.normal[2]; g.offset = q.offset;
            g.sx = q.start[0]; g.sy = q.start[1]; g.sz = q.start[2]; g.area = q.area;
            g.ux = q.u[0]; g.uy = q.u[1]; g.uz = q.u[2]; g._p0 = 0;
            g.vx = q.v[0]; g.vy = q.v[1]; g.vz = q.v[2]; g._p1 = 0;
            g.wx = q.w[0]; g.wy = q.w[1]; g.wz = q.w[2]; g._p2 = 0;
            g.ax = g.vy * g.wz - g.vz * g.wy; g.ay = g.vz * g.wx - g.vx * g.wz; g.az = g.vx * g.wy - g.vy * g.wx;  // v x w
            g.bx = g.wy * g.uz - g.wz * g.uy; g.by = g.wz * g.ux - g.wx * g.uz; g.bz = g.wx * g.uy - g.wy * g.ux;  // w x u
            out.quads[i] = g;
        }
        return true;
    }

    bool compile_lights() {
        out.has_lights = false;
        if (sc->lights == WRT_NONE) return true;
        if (!check_entity(sc->lights, "lights")) return false;
        const wrt_entity& L = sc->entities[sc->lights];
        auto add = [&](uint32_t id) -> bool {
            if (!check_entity(id, "lights child")) return false;
            const wrt_entity& e = sc->entities[id];
            Light l;
            l.index = 0;
            if (e.kind == WRT_ENT_SPHERE) {
                if (e.a >= sc->n_spheres) return fail(WRT_E_INVALID, "light sphere index out of range");
                if (sc->spheres[e.a].is_moving) return fail(WRT_E_INVALID, "moving spheres cannot be lights (entity.zig:627 asserts)");
                l.kind = WRT_ENT_SPHERE; l.index = e.a;
            } else if (e.kind == WRT_ENT_QUAD) {
                if (e.a >= sc->n_quads) return fail(WRT_E_INVALID, "light quad index out of range");
                l.kind = WRT_ENT_QUAD; l.index = e.a;
            } else if (e.kind == WRT_ENT_COLLECTION) {
                return fail(WRT_E_LIMIT, "nested light collections are not supported");
            } else {
                l.kind = e.kind;  // pdfValue 0, direction (1,0,0): entity.zig:47-65
            }
            out.lights.push_back(l);
            Box3 lb;
            lb.reset();  // stays empty (nothing passes) for the kinds whose pdfValue is 0
            if ((e.kind == WRT_ENT_SPHERE || e.kind == WRT_ENT_QUAD) && !tight_box(id, lb)) return false;
            out.light_boxes.push_back(padded(lb));
            return true;
        };
        if (L.kind == WRT_ENT_COLLECTION) {
            if (L.b == 0) return fail(WRT_E_INVALID, "light collection is empty (entity.zig:382 asserts len > 0)");
            if ((uint64_t)L.a + L.b > sc->n_children) return fail(WRT_E_INVALID, "light collection child range out of bounds");
            for (uint32_t k = 0; k < L.b; ++k)
                if (!add(sc->children[L.a + k])) return false;
        } else {
            if (!add(sc->lights)) return false;
        }
        out.has_lights = true;
        return true;
    }

    // The program WRT_CULL_TIGHT scans in packet form (closest_hit_packet): `ops` with the work removed that cannot pay for
    // itself when 32 rays share one program counter.  Tight boxes are conservative, so dropping a box test never changes a
    // result; WRT_CULL_REFERENCE must keep every node (its boxes are not conservative) and scans `ops`.
    //   * nodes whose box has >= 50 % of the surface area of their nearest kept ancestor (same coordinate frame; a
    //     translation keeps the extents, a rotation starts afresh): a subtree is skipped only when ALL lanes miss its box,
    //     which for so large a box practically never happens.  Cornell box: its BVH puts one wall in each half, so 7 of its
    //     8 bvh_node boxes are (nearly) the whole room.  The root is always kept (it is what stops rays that miss the scene).
    //   * a POP directly followed by another POP (the outer one recomputes the ray from the world ray anyway);
    //   * transform ops are flagged (z = 1) when no box test runs before the next transform op: the scan then skips the
    //     binary32 culler set-up (3 reciprocals + conversions) for that frame.  Safe on every execution path: a node that is
    //     reached by skipping a subtree sees the culler its subtree's root node was tested with, in the same frame.
    void prune_program() {
        const size_t n = out.ops.size();
        out.ops_pruned.clear();
        if (n > 1024) return;  // far beyond WRT_PACKET_MAX_OPS: the packet scan is never chosen for such a program
        struct Enclosing { uint32_t end; double area; };
        std::vector<Enclosing> stack;      // kept nodes around the current op, innermost last
        std::vector<size_t> frame_base;    // stack height at each open PUSH (ancestors below it live in another frame)
        std::vector<uint8_t> keep(n, 1);
        auto area_of = [&](uint32_t box) {
            const BoxTight& b = out.boxes_tight[box];
            const double dx = (double)b.max_x - b.min_x, dy = (double)b.max_y - b.min_y, dz = (double)b.max_z - b.min_z;
            if (!(dx >= 0.0 && dy >= 0.0 && dz >= 0.0)) return -1.0;  // empty / inverted: keep the test, it culls everything
            return 2.0 * (dx * dy + dy * dz + dz * dx);
        };
        auto is_xform = [](uint32_t kind) { return kind == OP_PUSH_TRANSLATE || kind == OP_PUSH_ROTATE_Y || kind == OP_POP; };
        auto is_node = [](uint32_t kind) { return kind == OP_NODE || kind == OP_NODE_TIGHT_ONLY; };
        for (size_t pc = 0; pc < n; ++pc) {
            const uint4 op = out.ops[pc];
            const size_t base = frame_base.empty() ? 0 : frame_base.back();
            while (stack.size() > base && stack.back().end <= pc) stack.pop_back();
            if (is_node(op.x)) {
                const double a = area_of(op.y);
                if (stack.size() > base && a >= 0.0 && stack.back().area > 0.0 && a >= 0.5 * stack.back().area) keep[pc] = 0;
                else stack.push_back({op.z, a});
            } else if (op.x == OP_PUSH_TRANSLATE) {
                // the enclosing bound keeps its extents under a translation: carry its area into the new frame
                const double a = (stack.size() > base) ? stack.back().area : -1.0;
                frame_base.push_back(stack.size());
                if (a > 0.0) stack.push_back({(uint32_t)n, a});
            } else if (op.x == OP_PUSH_ROTATE_Y) {
                frame_base.push_back(stack.size());
            } else if (op.x == OP_POP) {
                if (!frame_base.empty()) { stack.resize(frame_base.back()); frame_base.pop_back(); }
            }
        }
        // POP, POP -> POP
        for (size_t pc = 0; pc < n; ++pc) {
            if (!keep[pc] || out.ops[pc].x != OP_POP) continue;
            size_t nx = pc + 1;
            while (nx < n && !keep[nx]) ++nx;
            if (nx < n && out.ops[nx].x == OP_POP) keep[pc] = 0;
        }
        std::vector<uint32_t> renum(n + 1, 0);
        uint32_t next = 0;
        for (size_t pc = 0; pc < n; ++pc) { renum[pc] = next; next += keep[pc]; }
        renum[n] = next;
        out.ops_pruned.clear();
        out.ops_pruned.reserve(next);
        for (size_t pc = 0; pc < n; ++pc) {
            if (!keep[pc]) continue;
            uint4 op = out.ops[pc];
            if (is_node(op.x)) op.z = renum[op.z];
            out.ops_pruned.push_back(op);
        }
        bool changed = next != n;
        for (size_t pc = 0; pc < out.ops_pruned.size(); ++pc) {
            uint4& op = out.ops_pruned[pc];
            if (!is_xform(op.x)) continue;
            size_t nx = pc + 1;
            while (nx < out.ops_pruned.size() && !is_node(out.ops_pruned[nx].x) && !is_xform(out.ops_pruned[nx].x) &&
                   out.ops_pruned[nx].x != OP_END) ++nx;
            const bool culler_used = nx < out.ops_pruned.size() && is_node(out.ops_pruned[nx].x);
            op.z = culler_used ? 0u : 1u;
            changed = changed || !culler_used;
        }
        if (!changed) out.ops_pruned.clear();  // identical to `ops`: the packet scan uses that
    }

    int run() {
        if (!sc) { fail(WRT_E_INVALID, "scene is NULL"); return code; }
        if (sc->abi_version != WRT_ABI_VERSION) { fail(WRT_E_INVALID, "wrt_scene.abi_version mismatch"); return code; }
        if (sc->n_entities == 0 || sc->root >= sc->n_entities) { fail(WRT_E_INVALID, "scene root out of range"); return code; }
        if ((sc->n_entities && !sc->entities) || (sc->n_children && !sc->children) || (sc->n_spheres && !sc->spheres) ||
            (sc->n_quads && !sc->quads) || (sc->n_materials && !sc->materials) || (sc->n_textures && !sc->textures) ||
            (sc->n_images && !sc->images)) {
            fail(WRT_E_INVALID, "scene array pointer is NULL with a non-zero count");
            return code;
        }
        tight.resize(sc->n_entities);
        tight_state.assign(sc->n_entities, 0);
        prim_id.assign(sc->n_entities, WRT_NONE);
        on_stack.assign(sc->n_entities, 0);
        const bool trace = std::getenv("WRT_TRACE_BUILD") != nullptr;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        const auto t0 = now();
        if (!compile_materials() || !compile_geometry() || !compile_lights()) return code;
        const auto t1 = now();
        {  // size the output arrays once (a hint: shared subtrees are emitted once per use and may exceed it)
            size_t n_nodes = 0, n_xf = 0;
            for (uint32_t i = 0; i < sc->n_entities; ++i) {
                const uint32_t k = sc->entities[i].kind;
                n_nodes += (k == WRT_ENT_BVH_NODE);
                n_xf += (k == WRT_ENT_TRANSLATE || k == WRT_ENT_ROTATE_Y);
            }
            out.ops.reserve((size_t)sc->n_entities + 2 * n_xf + 16);
            out.boxes_ref.reserve(n_nodes + n_xf + 1);
            out.boxes_tight.reserve(n_nodes + n_xf + 1);
            out.nodes2.reserve(n_nodes + n_xf + 1);
            out.xforms.reserve(n_xf);
        }
        out.sphere_pc.assign(std::max<size_t>(sc->n_spheres, 1), WRT_NONE);
        out.quad_pc.assign(std::max<size_t>(sc->n_quads, 1), WRT_NONE);
        if (!emit(sc->root, WRT_NONE, 0)) return code;
        out.ops.push_back(make_uint4(OP_END, 0, 0, 0));
        const auto t2 = now();
        prune_program();
        if (trace) std::fprintf(stderr, "wrt trace: compile: records %.1f ms, program %.1f ms, prune %.1f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, now()));  // the trees of the ordered traversal are built after run(): build_trees_host / build_trees_device
        // transform chains in application order (outermost first), so the device needs no per-thread array
        out.xform_chains.assign(std::max<size_t>(out.xforms.size(), 1) * WRT_MAX_XFORM_DEPTH, WRT_NONE);
        for (size_t x = 0; x < out.xforms.size(); ++x) {
            uint32_t tmp[WRT_MAX_XFORM_DEPTH];
            int n = 0;
            for (uint32_t k = (uint32_t)x; k != WRT_NONE && n < WRT_MAX_XFORM_DEPTH; k = out.xforms[k].parent) tmp[n++] = k;
            for (int i = 0; i < n; ++i) out.xform_chains[x * WRT_MAX_XFORM_DEPTH + (size_t)i] = tmp[n - 1 - i];
        }
        if (out.boxes_ref.empty()) {  // keep the device pointers non-null
            BoxRef r = {0, 0, 0, 0};
            out.boxes_ref.push_back(r);
            Box3 e; e.reset();
            out.boxes_tight.push_back(padded(e));
            Node2 none;
            std::memset(&none, 0, sizeof none);
            none.l_desc = none.r_desc = WRT_NONE;
            out.nodes2.push_back(none);
        }
        return code;
    }
};

}  // namespace

// ---- ordered-traversal trees, host build ------------------------------------------------------------------------------
// The ordered traversal (wrt_device.cuh, Trav) only needs SOME binary tree over the leaf entities of each reference BVH:
// closest hit and tie rule are properties of the primitives and their DFS positions, not of the tree.  The reference splits
// at the median of a RANDOM axis (entity.zig:226-259); here each BVH is rebuilt over the same leaves with a binned
// surface-area heuristic on the tight boxes (wrt_treebuild.cuh), which roughly halves the nodes a ray visits on the
// 2^20-primitive scene.  `ops`, the reference boxes and everything WRT_CULL_REFERENCE reads keep the reference topology.
// WRT_REFERENCE_TREE=1 keeps it for the ordered traversal too.  The device builder (wrt_build.cu) produces the same bytes.
namespace {

struct HostTreeBuilder {
    CompiledScene& out;
    explicit HostTreeBuilder(CompiledScene& o) : out(o) {}

    // builds the subtree over items[lo, hi) (hi - lo >= 2) into record `rec`; its other hi - lo - 2 records are
    // nodes2[free, free + hi - lo - 2) (left subtree first), so the layout does not depend on which thread builds what and
    // large subtrees of the first levels are built concurrently.  Returns the subtree's depth in records.
    uint32_t build_sah(std::vector<TreeItem>& items, size_t lo, size_t hi, uint32_t rec, uint32_t level, uint32_t free) {
        double cmn[3], cmx[3];
        for (int k = 0; k < 3; ++k) { cmn[k] = INFINITY; cmx[k] = -INFINITY; }
        for (size_t i = lo; i < hi; ++i)
            for (int k = 0; k < 3; ++k) { const double c = tb_centroid(items[i], k); cmn[k] = fmin(cmn[k], c); cmx[k] = fmax(cmx[k], c); }
        double best_cost = INFINITY, best_base = 0.0, best_scale = 0.0;
        int best_axis = -1, best_bin = 0;
        if (level <= kTreeSahLevels) {
            for (int axis = 0; axis < 3; ++axis) {
                const double ext = cmx[axis] - cmn[axis];
                if (!(ext > 0.0)) continue;
                uint32_t bin_n[kTreeBins] = {};
                double bin_box[kTreeBins * 6];
                for (int b = 0; b < kTreeBins; ++b)
                    for (int k = 0; k < 3; ++k) { bin_box[b * 6 + k] = INFINITY; bin_box[b * 6 + 3 + k] = -INFINITY; }
                const double scale = (double)kTreeBins / ext;
                for (size_t i = lo; i < hi; ++i) {
                    const TreeItem& it = items[i];
                    const int b = tb_bin(tb_centroid(it, axis), cmn[axis], scale);
                    ++bin_n[b];
                    if (tb_valid(it.mn, it.mx))
                        for (int k = 0; k < 3; ++k) { bin_box[b * 6 + k] = fmin(bin_box[b * 6 + k], it.mn[k]); bin_box[b * 6 + 3 + k] = fmax(bin_box[b * 6 + 3 + k], it.mx[k]); }
                }
                double cost;
                int bin;
                tb_sweep_axis(bin_n, bin_box, cost, bin);
                if (bin >= 0 && cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = bin; best_base = cmn[axis]; best_scale = scale; }
            }
        }
        size_t mid;
        if (best_axis >= 0) {
            auto it = std::stable_partition(items.begin() + (ptrdiff_t)lo, items.begin() + (ptrdiff_t)hi, [&](const TreeItem& x) {
                return tb_bin(tb_centroid(x, best_axis), best_base, best_scale) <= best_bin;
            });
            mid = (size_t)(it - items.begin());
        } else {
            mid = lo + (hi - lo) / 2;  // no plane separates the centroids, or the depth cap: halve the current order
        }
        Node2 n;
        std::memset(&n, 0, sizeof n);
        const TreeChildren ch = tb_children((uint32_t)lo, (uint32_t)mid, (uint32_t)hi, free);
        for (int side = 0; side < 2; ++side) {
            const size_t a = side == 0 ? lo : mid, b = side == 0 ? mid : hi;
            double mn[3], mx[3];
            for (int k = 0; k < 3; ++k) { mn[k] = INFINITY; mx[k] = -INFINITY; }
            for (size_t i = a; i < b; ++i)
                if (tb_valid(items[i].mn, items[i].mx))
                    for (int k = 0; k < 3; ++k) { mn[k] = fmin(mn[k], items[i].mn[k]); mx[k] = fmax(mx[k], items[i].mx[k]); }
            float flo[3], fhi[3];
            tb_padded(mn, mx, flo, fhi);
            if (b - a == 1) tb_set_child(n, side, flo, fhi, items[a].start, items[a].end);
            else tb_set_child(n, side, flo, fhi, 0x80000000u | ch.rec[side], 0);
        }
        out.nodes2[rec] = n;
        uint32_t depth_l = 0, depth_r = 0;
        const bool fork = level <= 4 && ch.rec[0] != WRT_NONE && ch.rec[1] != WRT_NONE && (mid - lo) >= 8192 && (hi - mid) >= 8192;
        bool forked = false;
        if (fork) {
            std::future<uint32_t> left;
            try {
                left = std::async(std::launch::async, [&] { return build_sah(items, lo, mid, ch.rec[0], level + 1, ch.free[0]); });
                forked = true;
            } catch (const std::system_error&) {  // no thread to be had: build this level serially
                forked = false;
            }
            if (forked) {
                depth_r = build_sah(items, mid, hi, ch.rec[1], level + 1, ch.free[1]);
                depth_l = left.get();
            }
        }
        if (!forked) {
            if (ch.rec[0] != WRT_NONE) depth_l = build_sah(items, lo, mid, ch.rec[0], level + 1, ch.free[0]);
            if (ch.rec[1] != WRT_NONE) depth_r = build_sah(items, mid, hi, ch.rec[1], level + 1, ch.free[1]);
        }
        return 1 + std::max(depth_l, depth_r);
    }
    void rebuild_trees() {
        for (TreeInput& r : out.tree_inputs) {
            if (r.items.size() < 2) continue;  // a single leaf: the reference's record is already minimal
            const uint32_t free = (uint32_t)out.nodes2.size();
            Node2 none;
            std::memset(&none, 0, sizeof none);
            none.l_desc = none.r_desc = WRT_NONE;
            out.nodes2.resize(out.nodes2.size() + r.items.size() - 2, none);  // a tree over n leaves has n - 1 records, one is r.record
            const uint32_t depth = build_sah(r.items, 0, r.items.size(), r.record, 1, free);
            out.max_nesting = std::max(out.max_nesting, r.nest + depth);
        }
    }

    // Four-wide records: every tree of nodes2 (SAH-rebuilt or reference topology) collapsed top-down, breadth first
    // (wrt_treebuild.cuh): half the dependent record fetches per ray, one 128-byte line each.
    void build_nodes4() {
        out.nodes4.clear();
        out.root4.assign(out.nodes2.size(), WRT_NONE);
        struct Work { uint32_t rec2, rec4; };
        std::vector<Work> level, next_level;
        for (const TreeInput& r : out.tree_inputs) {
            out.root4[r.record] = (uint32_t)out.nodes4.size();
            out.nodes4.emplace_back();
            level.assign(1, Work{r.record, out.root4[r.record]});
            while (!level.empty()) {
                next_level.clear();
                for (const Work& w : level) {
                    TreeChild4 ch[4];
                    const int n = tb_widen(out.nodes2.data(), w.rec2, ch);
                    const uint32_t first_child = (uint32_t)out.nodes4.size();
                    for (int i = 0; i < n; ++i)
                        if (ch[i].desc & 0x80000000u) {
                            next_level.push_back(Work{ch[i].desc & 0x7FFFFFFFu, (uint32_t)out.nodes4.size()});
                            out.nodes4.emplace_back();
                        }
                    out.nodes4[w.rec4] = tb_node4(ch, n, first_child, out.ops.data());
                }
                level.swap(next_level);
            }
        }
    }
};

void finish_trees(CompiledScene& out) {
    if (out.nodes4.empty()) {  // keep the device pointer non-null
        Node4 n;
        std::memset(&n, 0, sizeof n);
        for (int i = 0; i < 4; ++i) n.desc[i] = WRT_NONE;
        out.nodes4.push_back(n);
    }
    const char* force_wide = std::getenv("WRT_WIDE_TREE");  // A/B switch: 0 / 1 overrides the size rule
    out.use_wide = force_wide ? (force_wide[0] == '1') : (out.nodes2.size() >= WRT_WIDE_TREE_MIN_RECORDS);
    out.stack_depth = ordered_stack_depth(out);
    // compact stack entries (wrt_device.cuh: TravCompactStack): one tree whose root op spans the program, no transforms, every
    // leaf of the four-wide records a single primitive, every primitive record tested by exactly one op
    out.compact_ok = false;
    if (out.use_wide && out.prim_pc_unique && out.xforms.empty() && out.tree_inputs.size() == 1 && out.ops.size() >= 2 &&
        out.ops[0].x == OP_NODE && out.ops[0].z == out.ops.size() - 1 && !out.has_moving) {
        bool ok = true;
        for (const Node4& n : out.nodes4)
            for (int i = 0; i < 4 && ok; ++i)
                if (n.desc[i] != WRT_NONE && !(n.desc[i] & 0x80000000u) && !(n.end[i] & WRT_LEAF_PRIM)) ok = false;
        out.compact_ok = ok;
    }
    // ... and their records in quantised form (Node4Q), first levels on the host threads
    out.nodes4q.clear();
    if (out.compact_ok) {
        out.nodes4q.resize(out.nodes4.size());
        const size_t n = out.nodes4.size();
        const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        const unsigned parts = n >= 65536 ? hw : 1u;
        std::vector<char> ok(parts, 1);
        auto work = [&](unsigned part) {
            const size_t lo = n * part / parts, hi = n * (part + 1) / parts;
            for (size_t i = lo; i < hi; ++i)
                if (!tb_node4q(out.nodes4[i], out.nodes4q[i])) { ok[part] = 0; return; }
        };
        std::vector<std::thread> threads;
        try {
            for (unsigned part = 1; part < parts; ++part) threads.emplace_back(work, part);
        } catch (const std::system_error&) {  // no thread to be had: the parts not started run here
            for (unsigned part = (unsigned)threads.size() + 1; part < parts; ++part) work(part);
        }
        work(0);
        for (auto& th : threads) th.join();
        for (char c : ok)
            if (!c) out.nodes4q.clear();  // an axis that cannot be expressed: the scene keeps walking nodes4
    }
    out.trees_built = true;
}

}  // namespace

bool keep_reference_trees() {
    const char* keep = std::getenv("WRT_REFERENCE_TREE");
    return keep && keep[0] == '1';
}

void build_trees_host(CompiledScene& out) {
    HostTreeBuilder b(out);
    if (!keep_reference_trees()) b.rebuild_trees();
    b.build_nodes4();
    finish_trees(out);
}

void finish_trees_after_device_build(CompiledScene& out) { finish_trees(out); }

namespace {

bool check_program(const std::vector<uint4>& ops, const char* what, std::string& err) {
    const size_t n = ops.size();
    if (n == 0 || ops[n - 1].x != OP_END) { err = std::string(what) + ": program does not end in OP_END"; return false; }
    int depth = 0;
    for (size_t pc = 0; pc < n; ++pc) {
        const uint4 op = ops[pc];
        if (op.x == OP_NODE || op.x == OP_NODE_TIGHT_ONLY) {
            if (!(op.z > pc && op.z <= n - 1)) { err = std::string(what) + ": skip link out of range at op " + std::to_string(pc); return false; }
        } else if (op.x == OP_PUSH_TRANSLATE || op.x == OP_PUSH_ROTATE_Y) {
            ++depth;
        } else if (op.x == OP_POP) {
            if (--depth < 0) { err = std::string(what) + ": unbalanced POP at op " + std::to_string(pc); return false; }
        } else if (op.x == OP_END && pc != n - 1) {
            err = std::string(what) + ": OP_END in the middle"; return false;
        } else if (op.x > OP_NODE_TIGHT_ONLY) {
            err = std::string(what) + ": unknown op kind"; return false;
        }
    }
    // a POP may have been fused away in the packet program: the remaining ones must still close every frame they leave
    return true;
}

}  // namespace

bool check_compiled_scene(const CompiledScene& cs, uint32_t& tree_depth, std::string& err) {
    tree_depth = 0;
    if (!check_program(cs.ops, "ops", err)) return false;
    const size_t n = cs.ops.size();
    if (!cs.ops_pruned.empty()) {
        // the packet program holds the same primitive and transform-entry ops in the same order; nodes and POPs may be fewer
        size_t j = 0;
        for (size_t pc = 0; pc < n; ++pc) {
            const uint4 op = cs.ops[pc];
            if (op.x != OP_SPHERE && op.x != OP_QUAD && op.x != OP_PUSH_TRANSLATE && op.x != OP_PUSH_ROTATE_Y) continue;
            while (j < cs.ops_pruned.size() && cs.ops_pruned[j].x != OP_SPHERE && cs.ops_pruned[j].x != OP_QUAD &&
                   cs.ops_pruned[j].x != OP_PUSH_TRANSLATE && cs.ops_pruned[j].x != OP_PUSH_ROTATE_Y) ++j;
            if (j == cs.ops_pruned.size() || cs.ops_pruned[j].x != op.x || cs.ops_pruned[j].y != op.y ||
                ((op.x == OP_SPHERE || op.x == OP_QUAD) && (cs.ops_pruned[j].z != op.z || cs.ops_pruned[j].w != op.w))) {
                err = "packet program differs from ops at op " + std::to_string(pc); return false;
            }
            ++j;
        }
        const size_t np = cs.ops_pruned.size();
        if (np == 0 || cs.ops_pruned[np - 1].x != OP_END) { err = "packet program does not end in OP_END"; return false; }
        for (size_t pc = 0; pc < np; ++pc) {
            const uint4 op = cs.ops_pruned[pc];
            if ((op.x == OP_NODE || op.x == OP_NODE_TIGHT_ONLY) && !(op.z > pc && op.z <= np - 1)) {
                err = "packet program: skip link out of range at op " + std::to_string(pc); return false;
            }
        }
    }
    // compact form (TravCompactStack / Node4Q): the primitive -> op tables invert the program, and every quantised box contains
    // the binary32 box it stands for, with the child words of the record it was made from
    if (cs.compact_ok) {
        for (size_t pc = 0; pc < n; ++pc) {
            const uint4 op = cs.ops[pc];
            if (op.x == OP_SPHERE && (op.y >= cs.sphere_pc.size() || cs.sphere_pc[op.y] != pc)) { err = "sphere_pc does not invert the program at op " + std::to_string(pc); return false; }
            if (op.x == OP_QUAD && (op.y >= cs.quad_pc.size() || cs.quad_pc[op.y] != pc)) { err = "quad_pc does not invert the program at op " + std::to_string(pc); return false; }
        }
        if (!cs.nodes4q.empty()) {
            if (cs.nodes4q.size() != cs.nodes4.size()) { err = "nodes4q: record count differs from nodes4"; return false; }
            for (size_t r = 0; r < cs.nodes4.size(); ++r) {
                const Node4& a = cs.nodes4[r];
                const Node4Q& q = cs.nodes4q[r];
                const float* lo[3] = {a.lox, a.loy, a.loz};
                const float* hi[3] = {a.hix, a.hiy, a.hiz};
                const double o[3] = {q.ox, q.oy, q.oz};
                for (int i = 0; i < 4; ++i) {
                    const uint32_t want = a.desc[i] == WRT_NONE ? WRT_NONE : ((a.desc[i] & 0x80000000u) ? a.desc[i] : (a.end[i] & 0x7FFFFFFFu));
                    if (q.word[i] != want) { err = "nodes4q: child word differs at record " + std::to_string(r); return false; }
                    if (a.desc[i] == WRT_NONE) continue;
                    for (int k = 0; k < 3; ++k) {
                        const double step = std::ldexp(1.0, (int)((q.exps >> (8 * k)) & 0xFFu) - 127);
                        const double dlo = o[k] + (double)((q.qlo[k] >> (8 * i)) & 0xFFu) * step, dhi = o[k] + (double)((q.qhi[k] >> (8 * i)) & 0xFFu) * step;
                        if (!(dlo <= (double)lo[k][i] && dhi >= (double)hi[k][i])) { err = "nodes4q: decoded box does not contain the child's box at record " + std::to_string(r); return false; }
                        if (!(dlo > (double)lo[k][i] - 1.0001 * step && dhi < (double)hi[k][i] + 1.0001 * step)) { err = "nodes4q: decoded box is more than one step loose at record " + std::to_string(r); return false; }
                    }
                }
            }
        }
    }
    // ordered-traversal trees: from every bvh root (an OP_NODE met inside a leaf range, plus op 0) the records must reach each
    // primitive op of [pc + 1, skip) exactly once (nested trees are entered through their own root op) — checked for the
    // child-pair records (nodes2) and for the four-wide records the traversal walks (nodes4)
    std::vector<uint8_t> seen(n, 0);
    for (int wide = 0; wide < 2; ++wide) {
        std::fill(seen.begin(), seen.end(), 0);
        const size_t n_records = wide ? cs.nodes4.size() : cs.nodes2.size();
        std::vector<std::pair<uint32_t, uint32_t>> ranges;  // leaf op ranges still to scan for nested roots
        ranges.push_back({0u, (uint32_t)n - 1});
        while (!ranges.empty()) {
            auto [lo, hi] = ranges.back();
            ranges.pop_back();
            for (uint32_t pc = lo; pc < hi;) {
                const uint4 op = cs.ops[pc];
                if (op.x != OP_NODE) {
                    if (op.x == OP_SPHERE || op.x == OP_QUAD) {
                        if (seen[pc]++) { err = "primitive op " + std::to_string(pc) + " reached twice"; return false; }
                    }
                    ++pc;
                    continue;
                }
                // walk this tree's records
                struct Item { uint32_t rec, depth; };
                std::vector<Item> st;
                uint32_t root = op.y;
                if (wide) {
                    if (op.y >= cs.root4.size() || cs.root4[op.y] == WRT_NONE) { err = "bvh root without a four-wide record at op " + std::to_string(pc); return false; }
                    root = cs.root4[op.y];
                }
                st.push_back({root, 1u});
                size_t visited = 0;
                while (!st.empty()) {
                    const Item it = st.back();
                    st.pop_back();
                    if (it.rec >= n_records) { err = "tree record index out of range"; return false; }
                    if (++visited > n_records) { err = "tree records form a cycle"; return false; }
                    if (!wide) tree_depth = std::max(tree_depth, it.depth);
                    uint32_t desc[4] = {WRT_NONE, WRT_NONE, WRT_NONE, WRT_NONE}, end[4] = {0, 0, 0, 0};
                    if (wide) {
                        for (int k = 0; k < 4; ++k) {
                            desc[k] = cs.nodes4[it.rec].desc[k]; end[k] = cs.nodes4[it.rec].end[k];
                            if (desc[k] != WRT_NONE && !(desc[k] & 0x80000000u) && (end[k] & WRT_LEAF_PRIM)) {  // single-primitive leaf
                                const uint4 lop = desc[k] < n ? cs.ops[desc[k]] : make_uint4(OP_END, 0, 0, 0);
                                const uint32_t kind = (end[k] & WRT_LEAF_QUAD) ? OP_QUAD : OP_SPHERE;
                                if (lop.x != kind || lop.y != (end[k] & WRT_LEAF_INDEX)) { err = "single-primitive leaf does not match its op"; return false; }
                                end[k] = desc[k] + 1;
                            }
                        }
                    } else {
                        const Node2& r = cs.nodes2[it.rec];
                        desc[0] = r.l_desc; desc[1] = r.r_desc; end[0] = r.l_end; end[1] = r.r_end;
                        if (r.l_desc == WRT_NONE) { err = "tree record without a left child"; return false; }
                    }
                    for (int k = 0; k < 4; ++k) {
                        if (desc[k] == WRT_NONE) continue;
                        if (desc[k] & 0x80000000u) st.push_back({desc[k] & 0x7FFFFFFFu, it.depth + 1});
                        else {
                            if (!(desc[k] > pc && end[k] > desc[k] && end[k] <= op.z)) { err = "tree leaf range outside its bvh at op " + std::to_string(pc); return false; }
                            ranges.push_back({desc[k], end[k]});
                        }
                    }
                }
                pc = op.z;  // the tree covered [pc + 1, skip)
            }
        }
        for (size_t pc = 0; pc < n; ++pc)
            if ((cs.ops[pc].x == OP_SPHERE || cs.ops[pc].x == OP_QUAD) && seen[pc] != 1) {
                err = std::string("primitive op ") + std::to_string(pc) + " is not reachable through the " + (wide ? "four-wide" : "child-pair") + " records";
                return false;
            }
    }
    return true;
}

namespace {
uint32_t stack_need_range(const CompiledScene& cs, uint32_t lo, uint32_t hi);
uint32_t stack_need_record(const CompiledScene& cs, uint32_t rec, uint32_t guard) {
  if (cs.use_wide) {
    if (rec >= cs.nodes4.size() || guard > 4096) return 1u << 20;  // malformed: never fits
    const Node4& r = cs.nodes4[rec];
    uint32_t need = 0, n_children = 0;
    for (int k = 0; k < 4; ++k) {
        if (r.desc[k] == WRT_NONE) continue;
        ++n_children;
        const uint32_t leaf_end = (r.end[k] & WRT_LEAF_PRIM) ? r.desc[k] + 1 : r.end[k];
        const uint32_t c = (r.desc[k] & 0x80000000u) ? stack_need_record(cs, r.desc[k] & 0x7FFFFFFFu, guard + 1) : stack_need_range(cs, r.desc[k], leaf_end);
        need = std::max(need, c);
    }
    return need + (n_children ? n_children - 1 : 0u);  // the other children wait on the stack while one is descended
  } else {
    if (rec >= cs.nodes2.size() || guard > 4096) return 1u << 20;  // malformed: never fits
    const Node2& r = cs.nodes2[rec];
    uint32_t need = 0;
    const uint32_t desc[2] = {r.l_desc, r.r_desc}, end[2] = {r.l_end, r.r_end};
    for (int k = 0; k < 2; ++k) {
        if (desc[k] == WRT_NONE) continue;
        const uint32_t c = (desc[k] & 0x80000000u) ? stack_need_record(cs, desc[k] & 0x7FFFFFFFu, guard + 1) : stack_need_range(cs, desc[k], end[k]);
        need = std::max(need, c);
    }
    const bool two = r.l_desc != WRT_NONE && r.r_desc != WRT_NONE;
    return need + (two ? 1u : 0u);  // the other child waits on the stack while this one is descended
  }
}
uint32_t stack_need_range(const CompiledScene& cs, uint32_t lo, uint32_t hi) {
    uint32_t need = 0;
    for (uint32_t pc = lo; pc < hi && pc < cs.ops.size();) {
        const uint4 op = cs.ops[pc];
        if (op.x == OP_NODE) {  // nested tree: the rest of the range waits while it is descended
            const uint32_t root = cs.use_wide ? (op.y < cs.root4.size() ? cs.root4[op.y] : WRT_NONE) : op.y;
            need = std::max(need, 1u + stack_need_record(cs, root, 0));
            pc = op.z > pc ? op.z : pc + 1;
        } else {
            ++pc;
        }
    }
    return need;
}
}  // namespace

uint32_t ordered_stack_depth(const CompiledScene& cs) {
    if (cs.ops.empty()) return 0;
    return stack_need_range(cs, 0, (uint32_t)cs.ops.size() - 1);
}

int compile_scene(const wrt_scene* scene, CompiledScene& out, std::string& err, bool defer_trees) {
    // nothing may unwind through the C ABI (a Zig or C caller): allocation failures of the vectors and std::system_error of
    // the concurrent SAH build come back as codes
    try {
        out = CompiledScene();
        Compiler c(scene, out, err);
        const int rc = c.run();
        if (rc == WRT_OK && !defer_trees) build_trees_host(out);
        return rc;
    } catch (const std::bad_alloc&) {
        err = "out of host memory while compiling the scene";
        return WRT_E_NOMEM;
    } catch (const std::exception& e) {
        err = std::string("scene compilation failed: ") + e.what();
        return WRT_E_INVALID;
    }
}

}  // namespace wrt
