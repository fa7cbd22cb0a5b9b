"""Builds a swgn_graph that represents a raw block-sparse linear least-squares problem
min |A x - b|^2 (the form of CERES/internal/ceres/linear_least_squares_problems.cc) so that the
reference's known-answer fixtures can be pushed through the C ABI: every row block becomes one
dense linear factor (MarginalizationFactor form r = r0 + J0 (x - x0)) zero-padded to a square J0,
with x = x0 = 0 and r0 = b, so residual = b and Jacobian = A at the evaluation point."""
import ctypes as C

import numpy as np

import swgn

i32, i64, f64 = C.c_int32, C.c_int64, C.c_double


class LinearGraph:
    def __init__(self, p):
        cs = list(p["col_sizes"])
        nb = len(cs)
        ne = p["num_eliminate_blocks"]
        self.keep = []
        g = swgn.Graph()
        self.block_size = np.array(cs, np.int32)
        self.block_manifold = np.zeros(nb, np.int32)
        self.block_const = np.zeros(nb, np.int32)
        self.block_group = np.array([0 if i < ne else i for i in range(nb)], np.int32)
        offs = np.concatenate([[0], np.cumsum(cs)]).astype(np.int32)
        self.block_offset = offs[:-1].copy()
        self.state = np.zeros(int(offs[-1]))
        prior_n, blk_begin, blocks, blk_idx = [], [0], [], []
        x0_begin, J_begin, r_begin = [], [], []
        x0, Jv, r0 = [], [], []
        v = 0
        row0 = 0
        for i, rs in enumerate(p["row_sizes"]):
            cells = p["cell_col"][p["row_ptr"][i]:p["row_ptr"][i + 1]]
            width = sum(cs[c] for c in cells)
            n = max(rs, width)
            J0 = np.zeros((n, n))
            col = 0
            x0_begin.append(len(x0))
            for c in cells:
                w = cs[c]
                J0[:rs, col:col + w] = np.array(p["values"][v:v + rs * w]).reshape(rs, w)
                v += rs * w
                blocks.append(c)
                blk_idx.append(col)
                x0 += [0.0] * w
                col += w
            rr = np.zeros(n)
            rr[:rs] = p["b"][row0:row0 + rs]
            row0 += rs
            prior_n.append(n)
            blk_begin.append(len(blocks))
            J_begin.append(len(Jv))
            Jv += list(J0.ravel())
            r_begin.append(len(r0))
            r0 += list(rr)
        self.prior_n = np.array(prior_n, np.int32)
        self.blk_begin = np.array(blk_begin, np.int32)
        self.blocks = np.array(blocks, np.int32)
        self.blk_idx = np.array(blk_idx, np.int32)
        self.x0_begin = np.array(x0_begin, np.int64)
        self.J_begin = np.array(J_begin, np.int64)
        self.r_begin = np.array(r_begin, np.int64)
        self.x0 = np.array(x0 + [0.0])
        self.Jv = np.array(Jv)
        self.r0 = np.array(r0)
        ip = lambda a: a.ctypes.data_as(C.POINTER(i32))
        lp = lambda a: a.ctypes.data_as(C.POINTER(i64))
        dp = lambda a: a.ctypes.data_as(C.POINTER(f64))
        g.n_blocks = nb
        g.block_size, g.block_manifold, g.block_const = ip(self.block_size), ip(self.block_manifold), ip(self.block_const)
        g.block_group, g.block_offset = ip(self.block_group), ip(self.block_offset)
        g.n_state = len(self.state)
        g.state = dp(self.state)
        g.proj_cauchy_a = 0.0
        g.n_prior = len(prior_n)
        g.prior_n, g.prior_blk_begin, g.prior_blocks, g.prior_blk_idx = ip(self.prior_n), ip(self.blk_begin), ip(self.blocks), ip(self.blk_idx)
        g.prior_x0_begin, g.prior_x0 = lp(self.x0_begin), dp(self.x0)
        g.prior_J_begin, g.prior_J = lp(self.J_begin), dp(self.Jv)
        g.prior_r_begin, g.prior_r0 = lp(self.r_begin), dp(self.r0)
        self.graph = g
        self.graph_p = C.pointer(g)
        self.n_cols = int(offs[-1])
        self.n_e = int(offs[ne])
        # rows of the padded problem (zero rows added) keep |Ax - b| unchanged
