"""Host-side mirror of the reference's mode-decision cost functions on top of the library's prediction primitives
(include/me_cuda.h: me_cuda_luma_prediction, me_cuda_pred_cost8x8).

Same names, argument meaning and return values as the reference (lencod/src/mv-search.c):
  GetSkipCostMB     :1136-1181   GetDirectCost8x8  :1273-1325   BIDPartitionCost  :1052-1128
The encoder state these functions read from globals (img->all_mv, direct_pdir, wbp_weight, ...) arrives as one PRED_BLOCK per
4x4 block of the macroblock, blocks[4 * by + bx]: what LumaPrediction (macroblock.c:832) is asked to do for that block.  The
binding of the real encoder (shim/jm_shim.c) fills the same records from the encoder's arrays.  Pixels never touch the host: the
library fetches the prediction from the planes in HBM and returns the costs.
"""
import numpy as np

INT_MAX = 2147483647
PRED_BLOCK_DTYPE = np.dtype([("pix_x", "<i2"), ("pix_y", "<i2"), ("p_dir", "i1"), ("ref", "i1", (2,)), ("apply_weights", "i1"),
                             ("mv", "<i2", (2, 2)), ("weight", "<i2", (2,)), ("offset", "<i2"), ("log_denom", "<i2")])
assert PRED_BLOCK_DTYPE.itemsize == 24
BLC_SIZE = {1: (16, 16), 2: (16, 8), 3: (8, 16), 4: (8, 8), 5: (8, 4), 6: (4, 8), 7: (4, 4)}       # input->blc_size, configfile.c:805-841


def mvbits(d):
    """mvbits[d] (mv-search.c:333-341)"""
    d = abs(int(d))
    return 1 if d == 0 else 2 * d.bit_length() + 1


def region_blocks(blocks16, bx, by):
    """the four 4x4 blocks of the 8x8 region whose top-left 4x4 block is (bx, by), raster order"""
    return [blocks16[4 * (by + j) + bx + i] for j in (0, 1) for i in (0, 1)]


def GetSkipCostMB(me, blocks16, rdopt, transform8x8, metric=2):
    """mv-search.c:1136-1181: blocks16 describe the skip prediction (p_dir 0, reference 0, the skip vector)"""
    regs = np.array([b for blk in range(4) for b in region_blocks(blocks16, 2 * (blk & 1), 2 * (blk >> 1))], PRED_BLOCK_DTYPE)
    c4, c8 = me.pred_cost8x8(regs, metric)
    use8 = rdopt == 0 and transform8x8
    return int((c8 if use8 else c4).sum())


def GetDirectCost8x8(me, blocks16, block, rdopt, transform8x8, metric=2, cost8x8=0):
    """mv-search.c:1273-1325 -> (cost, cost8x8); a block with p_dir < 0: direct mode not allowed, both INT_MAX"""
    regs = np.array(region_blocks(blocks16, 2 * (block % 2), 2 * (block // 2)), PRED_BLOCK_DTYPE)
    c4, c8 = me.pred_cost8x8(regs, metric)
    if c4[0] == INT_MAX:
        return INT_MAX, INT_MAX
    if rdopt == 0 and transform8x8:
        cost8x8 += int(c8[0])
    return int(c4[0]), cost8x8


def BIDPartitionCost(me, blocks16, pmv16, blocktype, block8x8, lambda_factor, transform8x8, metric=2):
    """mv-search.c:1052-1128: blocks16 carry p_dir 2 and all_mv[v][h][list][ref][blocktype]; pmv16[4 * v + h][list] = img->pred_mv"""
    bx0 = [[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 2, 0, 0], [0, 2, 0, 2]]
    by0 = [[0, 0, 0, 0], [0, 0, 0, 0], [0, 2, 0, 0], [0, 0, 0, 0], [0, 0, 2, 2]]
    parttype = blocktype if blocktype < 4 else 4
    step_h0, step_v0 = BLC_SIZE[parttype][0] >> 2, BLC_SIZE[parttype][1] >> 2
    step_h, step_v = BLC_SIZE[blocktype][0] >> 2, BLC_SIZE[blocktype][1] >> 2
    bx, by = bx0[parttype][block8x8], by0[parttype][block8x8]
    bits = 0
    for v in range(by, by + step_v0, step_v):
        for h in range(bx, bx + step_h0, step_h):
            b = blocks16[4 * v + h]
            for l in (0, 1):
                bits += mvbits(int(b["mv"][l][0]) - int(pmv16[4 * v + h][l][0])) + mvbits(int(b["mv"][l][1]) - int(pmv16[4 * v + h][l][1]))
    mcost = (lambda_factor * bits) >> 16                      # WEIGHTED_COST, defines.h:125
    regs = np.array([b for ry in range(by, by + step_v0, 2) for rx in range(bx, bx + step_h0, 2) for b in region_blocks(blocks16, rx, ry)],
                    PRED_BLOCK_DTYPE)
    c4, c8 = me.pred_cost8x8(regs, metric)
    use8 = transform8x8 and blocktype <= 4
    return mcost + int((c8 if use8 else c4).sum())
