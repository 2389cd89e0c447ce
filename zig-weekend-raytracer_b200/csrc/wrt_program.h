// wrt_program.h — host-side "flattener": compiles the entity tree handed over the C ABI (wrt_scene) into the
// stack-less traversal program and the 16-byte aligned record arrays the kernels read.
#pragma once

#include <string>
#include <vector>

#include "wrt_device.cuh"
#include "wrt_treebuild.cuh"

namespace wrt {

// the leaf entities of one reference BVH: what the ordered traversal's tree over them is built from
struct TreeInput {
    uint32_t record;  // the root's child-pair record (= box index of the bvh_node op)
    uint32_t nest;    // nesting of the root in the program
    std::vector<TreeItem> items;
};

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
    std::vector<Node4Q> nodes4q;               // compact_ok scenes: nodes4 quantised (empty: not available)
    std::vector<uint32_t> sphere_pc, quad_pc;  // per primitive record: the op that tests it (WRT_NONE: unused record)
    bool prim_pc_unique = true;                // no primitive record is tested by two ops
    bool compact_ok = false;                   // the scene qualifies for compact stack entries (set with the trees)
    std::vector<TreeInput> tree_inputs;  // per reference BVH, program order; consumed (reordered) by the tree build
    bool trees_built = false;            // nodes2 / nodes4 / root4 / use_wide / stack_depth are final
    uint32_t n_prims = 0;
    uint32_t max_xform_depth = 0;
    uint32_t max_nesting = 0;  // deepest chain of bvh_node / instance ops
    uint32_t stack_depth = 0;  // exact worst-case stack use of the ordered traversal over nodes2 (ordered_stack_depth)
    uint32_t ref_boxes_loose = 0;  // bvh_node boxes of the reference that do not contain their subtree's box in x / y
    bool has_lights = false;
    bool has_moving = false;
};

// Returns WRT_OK or a WRT_E_* code with `err` set.  With defer_trees the trees of the ordered traversal are left to the
// caller: build_trees_host (what compile_scene runs otherwise) or build_trees_device, which write the same bytes.
int compile_scene(const wrt_scene* scene, CompiledScene& out, std::string& err, bool defer_trees = false);

// SAH rebuild of every reference BVH over its leaves + four-wide collapse on the host threads (wrt_program.cu).
void build_trees_host(CompiledScene& cs);
// The same build on a CUDA device (wrt_build.cu): level-synchronous binned SAH, then the breadth-first collapse; the
// records come back into `cs`.  `ms` = device time of the build (events), excluding the transfers.  WRT_OK or a code.
int build_trees_device(CompiledScene& cs, int device, std::string& err, double* build_ms);
// true when the size rule (or WRT_DEVICE_BUILD=0/1) asks for the device build of this scene's trees
bool want_device_build(const CompiledScene& cs);
// compile_scene + the tree build where it belongs for this scene (device `device` or the host threads); what
// wrt_upload_scene and wrt_group_upload_scene run.  `tree_ms`: time of the tree build, `on_device`: where it ran.
int compile_scene_for_device(const wrt_scene* scene, CompiledScene& out, std::string& err, int device, double* tree_ms, bool* on_device);
bool keep_reference_trees();                            // WRT_REFERENCE_TREE=1
void finish_trees_after_device_build(CompiledScene& cs);  // use_wide, stack depth, sentinels

// Worst-case number of live stack entries of closest_hit_ordered (wrt_device.cuh) over this scene's trees: a child-pair
// record defers at most one child while it descends the other, a bvh_node met inside a leaf range defers the rest of the
// range.  The traversal is only used when this fits WRT_STACK_DEPTH.
uint32_t ordered_stack_depth(const CompiledScene& cs);

// Structural self-check of a compiled scene (host only; wrt_check_scene, tests/test_program.py): skip links, transform
// nesting, the packet program against the full one, and that every ordered-traversal tree reaches each primitive op of
// its BVH exactly once.  Returns true or sets `err`.  `tree_depth` = deepest tree in records.
bool check_compiled_scene(const CompiledScene& cs, uint32_t& tree_depth, std::string& err);

}  // namespace wrt
