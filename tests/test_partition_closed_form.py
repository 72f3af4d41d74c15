"""CPU check of the one non-obvious step of the device BVH build (csrc/bvh_device.cuh): the reference's in-place Hoare
partition (Raytracer/bvh.cpp:26-51) equals a closed form that needs only two prefix sums --
    A = positions where the up-scan stops (!(e < p)), ascending;  B = positions where the down-scan stops (!(e > p)), descending;
    swap A[k] <-> B[k] for k < m = #{k : A[k] < B[k]};  split index = min(A[m], B[m-1])  (A[0] when m = 0, "count" if A is empty).
The sequential loop below is the reference's with its unguarded scans guarded (running off an end can only mean "leaf")."""
import numpy as np


def hoare_sequential(e, p):
    e = e.copy(); n = len(e); i = -1; j = n
    while True:
        i += 1
        while i < n and e[i] < p:
            i += 1
        j -= 1
        while j >= 0 and e[j] > p:
            j -= 1
        if i >= j:
            break
        e[i], e[j] = e[j], e[i]
    return e, (n if i >= n else i)


def hoare_closed_form(e, p):
    n = len(e)
    a_flag = ~(e < p); b_flag = ~(e > p)
    scan_a = np.concatenate([[0], np.cumsum(a_flag)]); scan_b = np.concatenate([[0], np.cumsum(b_flag)])   # exclusive prefix sums
    n_a, n_b = scan_a[n], scan_b[n]
    pos_a = np.zeros(n, np.int64); pos_b = np.zeros(n, np.int64)
    for i in range(n):                                     # k_scatter: every position writes itself to its rank
        if a_flag[i]:
            pos_a[scan_a[i]] = i
        if b_flag[i]:
            pos_b[n_b - 1 - scan_b[i]] = i
    perm = np.arange(n); m = 0
    for k in range(min(n_a, n_b)):                         # k_pair: independent per k
        if pos_a[k] < pos_b[k]:
            perm[pos_a[k]], perm[pos_b[k]] = pos_b[k], pos_a[k]
            m = max(m, k + 1)
    if m == 0:
        split = pos_a[0] if n_a > 0 else n
    else:
        split = min(pos_a[m] if n_a > m else 1 << 60, pos_b[m - 1])
    return e[perm], int(split)


def test_closed_form_equals_the_sequential_partition():
    rng = np.random.RandomState(1)
    for _ in range(20000):
        n = rng.randint(1, 40)
        e = rng.randint(0, 8, size=n).astype(np.float32)            # many duplicates, many entries equal to the pivot
        p = np.float32(rng.randint(-1, 9)) + (np.float32(0.5) if rng.rand() < 0.3 else np.float32(0))
        a, sa = hoare_sequential(e, p)
        b, sb = hoare_closed_form(e, p)
        assert sa == sb and np.array_equal(a, b), (e, p, a, sa, b, sb)


def test_depth_first_numbering_from_subtree_counts():
    """the other half: breadth-first construction + renumbering reproduces the recursion's allocation order (bvh.cpp:259-272)"""
    rng = np.random.RandomState(2)
    for _ in range(200):
        # random binary tree, built recursively with the reference's numbering ...
        order = {}; counter = [2]

        def build(depth):
            node = {"kids": None}
            if depth < 8 and rng.rand() < 0.7:
                left = counter[0]; counter[0] += 2
                node["first_child_index"] = left
                node["kids"] = (build(depth + 1), build(depth + 1))
            return node
        root = build(0)
        # ... and renumbered from inner-node counts, level by level
        def inner(nd):
            nd["inner"] = 0 if nd["kids"] is None else 1 + inner(nd["kids"][0]) + inner(nd["kids"][1])
            return nd["inner"]
        inner(root)
        level = [(root, 0)]                                 # (node, pre-order rank among inner nodes)
        while level:
            nxt = []
            for nd, rank in level:
                if nd["kids"] is None:
                    continue
                assert nd["first_child_index"] == 2 + 2 * rank
                l, r = nd["kids"]
                nxt += [(l, rank + 1), (r, rank + 1 + l["inner"])]
            level = nxt


def test_udiv_magic_is_exact():
    """make_sampler's slot -> (pixel, sample) and pixel -> (row, column) divisions use a host-computed multiplier
    (csrc/device_math.cuh udiv_magic, csrc/kernels.cuh set_batch_magic): m = min(2^32 // d, 2^32 - 1), q = (n*m) >> 32,
    one correction.  Restated here in Python and checked against // and % over edge cases and random operands."""
    import random
    rnd = random.Random(3)

    def magic(d):
        return min((1 << 32) // d, 0xFFFFFFFF)

    def udiv(n, d, m):
        q = (n * m) >> 32
        r = n - q * d
        if r >= d:
            q += 1
            r -= d
        return q, r

    ds = list(range(1, 2050)) + [rnd.randrange(1, 1 << 32) for _ in range(5000)] + [(1 << 31) - 1, 1 << 31, (1 << 32) - 1, 65535, 65536, 65537]
    for d in ds:
        m = magic(d)
        for n in (0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 32) - 1, 1 << 31, rnd.randrange(1 << 32), rnd.randrange(1 << 32)):
            if 0 <= n < (1 << 32):
                assert udiv(n, d, m) == (n // d, n % d), (n, d)
