// Re-layout of a reference-format binary BVH (bpt_bvh_node[], Raytracer/bvh.h:31-45) into the two-level pair
// records the device traverses (wide_bvh.h).  Pure re-arrangement: boxes, child order, split axes and leaf ranges are
// copied verbatim; it also validates the node array (caller-supplied BVHs reach this through bpt_create_mesh_with_bvh).
#include <string.h>
#include <stdlib.h>
#include <vector>

#include "host_scene.h"

namespace bpt {

namespace {

struct WideTask {
    uint32_t node;          // binary inner node X whose children pair opens the record
    uint32_t ref_pair;      // the WChild that refers to X: pairs[ref_pair].c[ref_slot]; 0xFFFFFFFF = WideBVH::root
    uint32_t ref_slot;
    uint32_t depth;         // depth of X below the tree root
};

struct WideBuilder {
    const bpt_bvh_node* nodes;
    uint32_t node_count, item_count;
    WideBVH* out;
    std::vector<uint8_t> seen;
    const char* error = nullptr;

    bool is_leaf(const bpt_bvh_node& n) const { return n.count != 0 || item_count == 0; }      // leaf <=> count != 0 (bvh.h:35)

    bool visit(uint32_t index) {
        if (index >= node_count) { error = "child index past the node array"; return false; }
        if (seen[index]) { error = "node reached twice (cycle or shared subtree)"; return false; }
        seen[index] = 1;
        return true;
    }

    void fill_box(WChild* c, const bpt_bvh_node& n) {
        memcpy(c->p, n.bv_p, 12);
        memcpy(c->r, n.bv_r, 12);
    }

    bool fill_leaf(WChild* c, const bpt_bvh_node& n, uint32_t depth) {
        uint32_t first = n.left_first, count = n.count;
        if ((uint64_t)first + count > item_count) { error = "leaf range past the item array"; return false; }
        c->aux = count;
        if (count >= 1 && count <= BPT_WREF_INLINE_COUNT_MAX && first <= BPT_WREF_MAX_INDEX) {
            c->ref = wref_leaf(count, first);
        } else {
            if (out->big_leaves.size() > BPT_WREF_MAX_INDEX) { error = "too many oversized leaves"; return false; }
            c->ref = wref_leaf(0, (uint32_t)out->big_leaves.size());
            out->big_leaves.push_back({first, count});
        }
        if (depth > out->depth) out->depth = depth;
        return true;
    }
};

} // namespace

// mode 0: record roots are the inner nodes at even depth (every pair stored once, 1.5x the binary array);
// mode 1: every inner node opens a record of its own and its children's pairs are duplicated into it (3x), so that every
//         step of a descent covers two levels (experiment knob BPT_WIDE_MODE=1).
int build_wide_bvh(const bpt_bvh_node* nodes, uint32_t node_count, uint32_t item_count, WideBVH* out, int mode) {
    out->pairs.clear(); out->big_leaves.clear(); out->depth = 0; out->valid = false; out->mode = mode;
    memset(&out->root, 0, sizeof(out->root));
    if (!nodes || node_count == 0) { set_error("BVH re-layout: empty node array"); return BPT_ERR_ARG; }
    WideBuilder b;
    b.nodes = nodes; b.node_count = node_count; b.item_count = item_count; b.out = out;
    b.seen.assign(node_count, 0);
    b.seen[0] = 1;
    b.fill_box(&out->root, nodes[0]);

    std::vector<WideTask> stack;
    if (b.is_leaf(nodes[0])) {
        if (!b.fill_leaf(&out->root, nodes[0], 0)) goto fail;
    } else {
        stack.push_back({0u, 0xFFFFFFFFu, 0u, 0u});
    }
    while (!stack.empty()) {
        WideTask t = stack.back();
        stack.pop_back();
        const bpt_bvh_node& X = nodes[t.node];
        if (X.split_axis > 2) { b.error = "split axis out of range"; goto fail; }
        size_t base = out->pairs.size();
        if (base + 3 > BPT_WREF_MAX_INDEX) { b.error = "more pairs than a 28-bit reference can address"; goto fail; }
        out->pairs.resize(base + 3);
        memset(&out->pairs[base], 0, 3*sizeof(WPair));
        {
            WChild* who = t.ref_pair == 0xFFFFFFFFu ? &out->root : &out->pairs[t.ref_pair].c[t.ref_slot];
            who->ref = wref_inner(X.split_axis, true, (uint32_t)base);
            who->aux = t.node;
        }
        // children are pushed in reverse so that the record of c0's first grandchild follows this one (depth-first order)
        WideTask pending[4]; int n_pending = 0;
        for (uint32_t k = 0; k < 2; ++k) {
            uint32_t ci = X.left_first + k;
            if (!b.visit(ci)) goto fail;
            const bpt_bvh_node& c = nodes[ci];
            b.fill_box(&out->pairs[base].c[k], c);
            if (b.is_leaf(c)) {
                if (!b.fill_leaf(&out->pairs[base].c[k], c, t.depth + 1)) goto fail;
                continue;
            }
            if (c.split_axis > 2) { b.error = "split axis out of range"; goto fail; }
            if (mode == 1) {
                // the child opens its own record later; this record only carries a COPY of its pair (filled when the
                // child's record exists: remember where)
                pending[n_pending++] = {ci, (uint32_t)base, k, t.depth + 1};
                continue;
            }
            out->pairs[base].c[k].ref = wref_inner(c.split_axis, false, (uint32_t)(base + 1 + k));
            out->pairs[base].c[k].aux = ci;
            for (uint32_t j = 0; j < 2; ++j) {
                uint32_t gi = c.left_first + j;
                if (!b.visit(gi)) goto fail;
                const bpt_bvh_node& g = nodes[gi];
                WChild* wg = &out->pairs[base + 1 + k].c[j];
                b.fill_box(wg, g);
                if (b.is_leaf(g)) { if (!b.fill_leaf(wg, g, t.depth + 2)) goto fail; }
                else pending[n_pending++] = {gi, (uint32_t)(base + 1 + k), j, t.depth + 2};
            }
        }
        for (int i = n_pending - 1; i >= 0; --i) stack.push_back(pending[i]);
    }
    if (mode == 1) {
        // second pass: copy each inner child's finished pair into its parent's record
        for (size_t base = 0; base < out->pairs.size(); base += 3) {
            for (uint32_t k = 0; k < 2; ++k) {
                uint32_t ref = out->pairs[base].c[k].ref;
                if (ref & BPT_WREF_LEAF) continue;
                out->pairs[base + 1 + k] = out->pairs[ref & BPT_WREF_INDEX_MASK];
            }
        }
    }
    out->valid = true;
    return BPT_OK;
fail:
    set_error("BVH re-layout: invalid node array (%s)", b.error ? b.error : "?");
    out->pairs.clear(); out->big_leaves.clear();
    return BPT_ERR_ARG;
}

int build_wide_bvh(HostBVH* bvh, uint32_t item_count) {
    int mode = 0;
    if (const char* e = getenv("BPT_WIDE_MODE")) mode = atoi(e) == 1 ? 1 : 0;
    return build_wide_bvh(bvh->nodes.data(), (uint32_t)bvh->nodes.size(), item_count, &bvh->wide, mode);
}

} // namespace bpt
