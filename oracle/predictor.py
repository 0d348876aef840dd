"""TEST INFRASTRUCTURE (checker, never used by the product): plain-Python restatement of the reference's motion-vector
prediction and of the loop PartitionMotionSearch runs around its block searches, for frame macroblocks (no MBAFF).

  mv_predictor()          SetMotionVectorPredictor, lencod/src/mv-search.c:87-281, with the neighbour rules of
                          getLuma4x4Neighbour / getNonAffNeighbour (lencod/src/mb_access.c:549, :143)
  partition_blocks()      sub-blocks of PartitionMotionSearch(blocktype, block8x8) in call order, mv-search.c:1387-1401, :1424-1430
  partition_motion_search the loop itself (mv-search.c:1404-1452) around a caller-supplied block search, including the write-back of
                          reference index and vector that the predictors of the following sub-blocks read (:1441-1449)

The motion state is the layout of me_mb_motion (include/me_cuda.h): ref_idx[list][row + 1][col + 1], mv[list][row + 1][col + 1][2]
for the 4x4 blocks of rows -1..3 and columns -1..4 around the macroblock, ref_idx -128 = block not available.

Parity: pinned indirectly -- the byte-identical encoder runs of tests/test_jm_dropin.py answer a drop-in call from a device
result only when the predictor the device derived equals the one the reference's own SetMotionVectorPredictor computed, and
count every mismatch; tests/test_gpu_parity.py::test_partition_search_launch checks the device against this file.
"""
UNAVAILABLE = -128
BSX = {1: 16, 2: 16, 3: 8, 4: 8, 5: 8, 6: 4, 7: 4}
BSY = {1: 16, 2: 8, 3: 16, 4: 8, 5: 4, 6: 8, 7: 4}


def _neighbour(ref_idx, x, y):
    """pixel (x, y) relative to the macroblock -> (row, col) in the state, or None when not available"""
    if x >= 16 and y >= 0:
        return None                                  # to the right of the macroblock: not coded yet (mb_access.c:171-174)
    col = (-1 if x < 0 else (4 if x >= 16 else x >> 2)) + 1
    row = (-1 if y < 0 else y >> 2) + 1
    return (row, col) if ref_idx[row][col] != UNAVAILABLE else None


def mv_predictor(ref_idx, mv, ref_frame, block_x, block_y, bsx, bsy):
    """ref_idx / mv: state of ONE list.  (block_x, block_y) in 4x4 units inside the macroblock.  Returns (pmv_x, pmv_y)."""
    mb_x, mb_y = 4 * block_x, 4 * block_y
    a = _neighbour(ref_idx, mb_x - 1, mb_y)
    b = _neighbour(ref_idx, mb_x, mb_y - 1)
    c = _neighbour(ref_idx, mb_x + bsx, mb_y - 1)
    d = _neighbour(ref_idx, mb_x - 1, mb_y - 1)
    if mb_y > 0:                                     # mv-search.c:111-131
        if mb_x < 8:
            if mb_y == 8:
                if bsx == 16:
                    c = None
            elif mb_x + bsx == 8:
                c = None
        elif mb_x + bsx == 16:
            c = None
    if c is None:
        c = d
    r_l = ref_idx[a[0]][a[1]] if a else -1
    r_u = ref_idx[b[0]][b[1]] if b else -1
    r_ur = ref_idx[c[0]][c[1]] if c else -1
    kind = "median"
    if r_l == ref_frame and r_u != ref_frame and r_ur != ref_frame:
        kind = "l"
    elif r_l != ref_frame and r_u == ref_frame and r_ur != ref_frame:
        kind = "u"
    elif r_l != ref_frame and r_u != ref_frame and r_ur == ref_frame:
        kind = "ur"
    if bsx == 8 and bsy == 16:                       # directional predictions, mv-search.c:187-212
        if mb_x == 0:
            if r_l == ref_frame:
                kind = "l"
        elif r_ur == ref_frame:
            kind = "ur"
    elif bsx == 16 and bsy == 8:
        if mb_y == 0:
            if r_u == ref_frame:
                kind = "u"
        elif r_l == ref_frame:
            kind = "l"
    out = []
    for hv in (0, 1):
        mv_a = int(mv[a[0]][a[1]][hv]) if a else 0
        mv_b = int(mv[b[0]][b[1]][hv]) if b else 0
        mv_c = int(mv[c[0]][c[1]][hv]) if c else 0
        if kind == "median":
            pv = mv_a if not (b or c) else mv_a + mv_b + mv_c - min(mv_a, mv_b, mv_c) - max(mv_a, mv_b, mv_c)
        else:
            pv = {"l": mv_a, "u": mv_b, "ur": mv_c}[kind]
        out.append(((pv + 32768) & 0xffff) - 32768)   # short pmv[2]
    return tuple(out)


def partition_blocks(blocktype, block8x8):
    parttype = blocktype if blocktype < 4 else 4
    psx = [4, 4, 4, 2, 2, 2, 1, 1]
    psy = [4, 4, 2, 4, 2, 1, 2, 1]
    bx0 = [[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 2, 0, 0], [0, 2, 0, 2]]
    by0 = [[0, 0, 0, 0], [0, 0, 0, 0], [0, 2, 0, 0], [0, 0, 0, 0], [0, 0, 2, 2]]
    by, bx = by0[parttype][block8x8], bx0[parttype][block8x8]
    return [(h, v) for v in range(by, by + psy[parttype], psy[blocktype]) for h in range(bx, bx + psx[parttype], psx[blocktype])]


def partition_motion_search(state_ref_idx, state_mv, blocktype, block8x8, list_sizes, block_search):
    """state_*: [list][5][6] / [list][5][6][2] nested lists (modified in place like enc_picture).  block_search(list, ref, h, v,
    blocktype, pmv) -> (mv_x, mv_y, cost) in quarter pel.  Returns {(list, ref, h, v): (pmv, mv, cost)}."""
    res = {}
    bsx, bsy = BSX[blocktype], BSY[blocktype]
    for lst, n_refs in enumerate(list_sizes):
        for ref in range(n_refs):
            for (h, v) in partition_blocks(blocktype, block8x8):
                pmv = mv_predictor(state_ref_idx[lst], state_mv[lst], ref, h, v, bsx, bsy)
                mvx, mvy, cost = block_search(lst, ref, h, v, blocktype, pmv)
                res[(lst, ref, h, v)] = (pmv, (mvx, mvy), cost)
                for j in range(v, v + bsy // 4):
                    for i in range(h, h + bsx // 4):
                        state_ref_idx[lst][j + 1][i + 1] = ref
                        state_mv[lst][j + 1][i + 1] = [mvx, mvy]
    return res
