"""CPU tests of the device BVH layout (csrc/wide_bvh.h / wide_bvh.cpp): the two-level pair records must be a pure
re-arrangement of the reference-format node array -- decoding them gives back every node's box bits, split axis, child
order and leaf range (the node array itself is memcmp-equal to the reference's, tests/test_host_parity.py)."""
import os

import numpy as np
import pytest

from buas_pathtracer_b200 import capi, scenes

LEAF, ROOT, MASK = capi.WREF_LEAF, capi.WREF_RECORD_ROOT, capi.WREF_INDEX_MASK


def _decode_and_compare(nodes, item_count, children, root, big, depth, expect_all_roots=False):
    """walk the binary tree and the records together; returns (pairs visited, record-root steps, deepest leaf)"""
    def same_box(ch, nd):
        return (np.array_equal(ch["bv_p"].view(np.uint32), nd["bv_p"].view(np.uint32)) and
                np.array_equal(ch["bv_r"].view(np.uint32), nd["bv_r"].view(np.uint32)))

    def leaf_range(ref):
        cnt, first = (int(ref) >> 28) & 7, int(ref) & MASK
        if cnt == 0:
            first, cnt = int(big[first][0]), int(big[first][1])
        return first, cnt

    seen_pairs = set()
    stats = dict(pairs=0, roots=0, deepest=0, leaves=0)
    stack = [(0, root, 0)]         # (binary node index, child record, depth)
    while stack:
        ni, ch, d = stack.pop()
        nd = nodes[ni]
        assert same_box(ch, nd), f"node {ni}: box bits differ"
        ref = int(ch["ref"])
        is_leaf = nd["count"] != 0 or item_count == 0
        assert bool(ref & LEAF) == bool(is_leaf), f"node {ni}: leaf flag"
        if is_leaf:
            first, cnt = leaf_range(ref)
            assert (first, cnt) == (int(nd["left_first"]), int(nd["count"])), f"node {ni}: leaf range"
            assert int(ch["aux"]) == cnt
            stats["leaves"] += 1
            stats["deepest"] = max(stats["deepest"], d)
            continue
        assert (ref >> 29) & 3 == int(nd["split_axis"]), f"node {ni}: split axis"
        pi = ref & MASK
        is_root = bool(ref & ROOT)
        assert is_root == (d % 2 == 0 or expect_all_roots), f"node {ni} at depth {d}: record-root flag"
        if is_root:
            assert pi % 3 == 0
            stats["roots"] += 1
        assert (pi, d) not in seen_pairs
        seen_pairs.add((pi, d))
        stats["pairs"] += 1
        lf = int(nd["left_first"])
        for k in (0, 1):
            c = children[2 * pi + k]
            stack.append((lf + k, c, d + 1))
            if is_root and not (int(c["ref"]) & LEAF):
                # the record carries the child's own pair right behind the node's: what the two-level step reads
                cn = nodes[lf + k]
                for j in (0, 1):
                    g = children[2 * (pi + 1 + k) + j]
                    assert same_box(g, nodes[int(cn["left_first"]) + j]), f"node {ni}: grandchild pair {k}/{j} box"
                    gn = nodes[int(cn["left_first"]) + j]
                    assert bool(int(g["ref"]) & LEAF) == bool(gn["count"] != 0)
                    if not expect_all_roots:
                        assert (int(c["ref"]) & MASK) == pi + 1 + k, "odd-level child must live inside its parent's record"
    assert stats["deepest"] == depth
    return stats


@pytest.mark.parametrize("level", [0, 1, 3, 5])
def test_wide_layout_decodes_to_the_binary_tree(bpt, level):
    s = bpt.Scene()
    scenes.c2_icosphere(s, 64, 36, level=level)
    nodes, idx, _ = s.mesh_bvh(0)
    children, root, big, depth = s.wide_bvh(0)
    st = _decode_and_compare(nodes, len(idx), children, root, big, depth)
    inner = int(np.count_nonzero(nodes["count"][2:] == 0)) + (1 if nodes["count"][0] == 0 and len(idx) else 0)
    assert st["pairs"] == inner
    assert len(children) == 6 * st["roots"]               # three pairs per record, nothing else stored
    # the TLAS of the same scene
    tn, ti = s.scene_bvh()
    tc, tr, tb, td = s.wide_bvh(-1)
    _decode_and_compare(tn, len(ti), tc, tr, tb, td)


def test_wide_layout_instances_and_midpoint_trees(bpt):
    s = bpt.Scene()
    scenes.c3_instances(s, 64, 36, level=2, grid=4, sky_size=(64, 32))
    tn, ti = s.scene_bvh()
    tc, tr, tb, td = s.wide_bvh(-1)
    st = _decode_and_compare(tn, len(ti), tc, tr, tb, td)
    assert st["leaves"] >= 4 and td >= 2
    # a skewed midpoint-split tree (deep, with leaves at every depth)
    rng = np.random.RandomState(3)
    n = 600
    base = np.cumsum(rng.rand(n) ** 6 * 4.0).astype(np.float32)
    tris = np.zeros((n, 9), np.float32)
    tris[:, 0] = base; tris[:, 3] = base + 0.1; tris[:, 6] = base
    tris[:, 4] = 0.1; tris[:, 8] = 0.1
    m = s.create_mesh(tris, method=capi.BVH_MIDPOINT_SPLIT)
    nodes, idx, _ = s.mesh_bvh(m)
    children, root, big, depth = s.wide_bvh(m)
    _decode_and_compare(nodes, len(idx), children, root, big, depth)


def test_wide_layout_big_leaves(bpt):
    """coincident centroids cannot be split (bvh.cpp:254): one forced leaf with more than 7 triangles -> side table"""
    s = bpt.Scene()
    tri = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    m = s.create_mesh(np.repeat(tri, 23, axis=0))
    nodes, idx, _ = s.mesh_bvh(m)
    children, root, big, depth = s.wide_bvh(m)
    assert len(big) >= 1 and int(big[:, 1].max()) > 7
    _decode_and_compare(nodes, len(idx), children, root, big, depth)


def test_wide_layout_full_duplication_mode(bpt):
    """BPT_WIDE_MODE=1 (experiment): every inner node opens a record; same tree"""
    os.environ["BPT_WIDE_MODE"] = "1"
    try:
        s = bpt.Scene()
        scenes.c2_icosphere(s, 64, 36, level=3)
        nodes, idx, _ = s.mesh_bvh(0)
        children, root, big, depth = s.wide_bvh(0)
    finally:
        del os.environ["BPT_WIDE_MODE"]
    st = _decode_and_compare(nodes, len(idx), children, root, big, depth, expect_all_roots=True)
    assert len(children) == 6 * st["pairs"]


def test_malformed_caller_bvh_is_rejected(bpt):
    """bpt_create_mesh_with_bvh validates the node array while re-laying it out (cycles, children / leaf ranges out of
    bounds) instead of letting the device traversal read out of bounds or spin"""
    s = bpt.Scene()
    scenes.c2_icosphere(s, 64, 36, level=2)
    nodes, idx, tris = s.mesh_bvh(0)
    pos = np.zeros((len(idx), 9), np.float32)
    pos[idx] = tris
    assert s.create_mesh_with_bvh(pos, nodes, idx) == 1
    inner = np.flatnonzero(nodes["count"] == 0)
    inner = inner[inner != 1]
    bad = nodes.copy(); bad["left_first"][inner[3]] = len(nodes) + 10          # child past the array
    with pytest.raises(bpt.BptError):
        s.create_mesh_with_bvh(pos, bad, idx)
    bad = nodes.copy(); bad["left_first"][inner[5]] = 2                         # back edge: node 2 reached twice
    with pytest.raises(bpt.BptError):
        s.create_mesh_with_bvh(pos, bad, idx)
    leaves = np.flatnonzero(nodes["count"] != 0)
    bad = nodes.copy(); bad["left_first"][leaves[0]] = len(idx) - 1; bad["count"][leaves[0]] = 4   # leaf range past the triangles
    with pytest.raises(bpt.BptError):
        s.create_mesh_with_bvh(pos, bad, idx)
    bad = nodes.copy(); bad["split_axis"][inner[2]] = 3
    with pytest.raises(bpt.BptError):
        s.create_mesh_with_bvh(pos, bad, idx)
    assert s.counts()["meshes"] == 2                                            # the rejected meshes left nothing behind
