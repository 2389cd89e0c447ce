// wrt_program.h — host-side "flattener": compiles the entity tree handed over the C ABI (wrt_scene) into the
// stack-less traversal program and the 16-byte aligned record arrays the kernels read.
#pragma once

#include <string>
#include <vector>

#include "wrt_device.cuh"

namespace wrt {

struct CompiledScene {
    std::vector<uint4> ops;
    std::vector<uint4> ops_pruned;  // `ops` minus the nodes that cannot cull (packet scan under WRT_CULL_TIGHT); empty = same as ops
    std::vector<BoxRef> boxes_ref;
    std::vector<BoxTight> boxes_tight;
    std::vector<Node2> nodes2;  // parallel to boxes_*: child-pair records of the bvh_node ops (ordered traversal)
    std::vector<Node4> nodes4;  // the same trees collapsed to four-wide records (what the ordered traversal walks)
    bool use_wide = false;        // the ordered traversal walks nodes4 (large trees) instead of nodes2
    std::vector<uint32_t> root4;  // per box index: Node4 record of the tree rooted at that bvh_node op, WRT_NONE otherwise
    std::vector<SphereGeom> spheres;
    std::vector<SphereAux> sphere_aux;  // empty unless a sphere moves
    std::vector<QuadGeom> quads;
    std::vector<Xform> xforms;
    std::vector<uint32_t> xform_chains;  // WRT_MAX_XFORM_DEPTH ids per xform: the chain root -> leaf, WRT_NONE padded
    std::vector<Material> materials;
    std::vector<Texture> textures;
    std::vector<Light> lights;
    std::vector<BoxTight> light_boxes;  // conservative binary32 box of every light (an empty box for kinds whose pdf is 0)
    uint32_t n_prims = 0;
    uint32_t max_xform_depth = 0;
    uint32_t max_nesting = 0;  // deepest chain of bvh_node / instance ops
    uint32_t stack_depth = 0;  // exact worst-case stack use of the ordered traversal over nodes2 (ordered_stack_depth)
    uint32_t ref_boxes_loose = 0;  // bvh_node boxes of the reference that do not contain their subtree's box in x / y
    bool has_lights = false;
    bool has_moving = false;
};

// Returns WRT_OK or a WRT_E_* code with `err` set.
int compile_scene(const wrt_scene* scene, CompiledScene& out, std::string& err);

// Worst-case number of live stack entries of closest_hit_ordered (wrt_device.cuh) over this scene's trees: a child-pair
// record defers at most one child while it descends the other, a bvh_node met inside a leaf range defers the rest of the
// range.  The traversal is only used when this fits WRT_STACK_DEPTH.
uint32_t ordered_stack_depth(const CompiledScene& cs);

// Structural self-check of a compiled scene (host only; wrt_check_scene, tests/test_program.py): skip links, transform
// nesting, the packet program against the full one, and that every ordered-traversal tree reaches each primitive op of
// its BVH exactly once.  Returns true or sets `err`.  `tree_depth` = deepest tree in records.
bool check_compiled_scene(const CompiledScene& cs, uint32_t& tree_depth, std::string& err);

}  // namespace wrt
