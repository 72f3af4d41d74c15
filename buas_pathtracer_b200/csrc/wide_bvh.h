// Device layout of the acceleration structures: the reference's binary BVH (Raytracer/bvh.h:31-45, built bit-identically
// by bvh_build.cpp / bvh_device.cuh) re-laid-out as compact two-level records that are read with 128-bit loads.
//
// The binary tree itself is untouched -- same nodes, same boxes (the reference's centre / half-extent floats, verbatim),
// same child order, same split axes, same leaf ranges -- so a traversal that replays the reference's visit order over it
// (trace.cuh) produces the reference's hit records.  What changes is where the nodes live:
//
//   * the unit is the SIBLING PAIR (64 bytes = 4 x 128 bit): both children of one inner node, each child as
//       {bv_p.x, bv_p.y, bv_p.z, bv_r.x | bv_r.y, bv_r.z, ref, aux}
//     with everything a visit of that child needs packed into `ref` (so a traversal-stack entry is 8 bytes: ref + entry
//     distance):
//       inner child:  bit 31 = 0 | bits 29-30 = the child's split axis | bit 28 = "record root" | bits 0-27 = index of
//                     the pair holding the child's own children
//       leaf child:   bit 31 = 1 | bits 28-30 = item count 1..7 (0: `first` indexes the big-leaf table, for the forced
//                     leaves of bvh.cpp:254,278 that hold more) | bits 0-27 = first item (leaf-order triangle / TLAS index)
//   * pairs are grouped into RECORDS of three consecutive pairs (192 bytes): the pair of a record-root node X, then the
//     pair of X's child 0, then the pair of X's child 1 (zero-filled when that child is a leaf).  Record roots are the
//     inner nodes at even depth below the tree root.  Entering a record root, the traversal knows from X's split axis and
//     the ray's direction sign which child is the near one BEFORE anything is loaded, so it fetches pair(X) and the near
//     child's pair together: one dependent memory round trip per TWO levels of descent instead of one per level.
//     Nodes at odd depth that are entered from the stack (a popped far child) find their pair inside their parent's
//     record.  Records are laid out in depth-first order like the reference's node array.
//
// This header is shared by the host builder (wide_bvh.cpp) and the device code; it has no CUDA types.
#pragma once
#include <stdint.h>

namespace bpt {

struct WChild {              // 32 bytes
    float    p[3];           // bv_p
    float    r[3];           // bv_r
    uint32_t ref;
    uint32_t aux;            // leaf: item count (also for big leaves); inner: index of the binary node (diagnostics)
};
struct WPair { WChild c[2]; };      // 64 bytes: children left_first, left_first + 1 of one inner node

struct WBigLeaf { uint32_t first, count; };

#define BPT_WREF_LEAF        0x80000000u
#define BPT_WREF_INDEX_MASK  0x0FFFFFFFu
#define BPT_WREF_RECORD_ROOT 0x10000000u
#define BPT_WREF_MAX_INDEX   0x0FFFFFFFu
#define BPT_WREF_INLINE_COUNT_MAX 7u

static inline uint32_t wref_inner(uint32_t axis, bool record_root, uint32_t pair_index) {
    return ((axis & 3u) << 29) | (record_root ? BPT_WREF_RECORD_ROOT : 0u) | (pair_index & BPT_WREF_INDEX_MASK);
}
static inline uint32_t wref_leaf(uint32_t inline_count, uint32_t first_or_big_index) {
    return BPT_WREF_LEAF | ((inline_count & 7u) << 28) | (first_or_big_index & BPT_WREF_INDEX_MASK);
}

} // namespace bpt
