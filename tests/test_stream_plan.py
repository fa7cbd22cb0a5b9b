"""Host-side check of the streamed Schur plan (csrc/plan_stream.cpp): the record packages the CUDA kernel
k_schur_stream interprets are interpreted here, in numpy, on the oracle's Jacobian, batch by batch exactly as the
kernel walks them (operand area, chunk factors, W blocks, MMA-term runs into compact accumulators, write-out through
the cell map), and the resulting reduced system is compared with the oracle's Schur complement.  No GPU involved:
this pins the planner; tests/test_gpu_parity.py pins the kernel."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
import swgn

# csrc/device_types.h IArr
I_COL_SIZE, I_COL_POS, I_TCOL = 1, 3, 5
I_ROW_RES, I_ROW_NRES, I_ROW_CELL = 6, 7, 8
I_CELL_COL, I_CELL_VAL = 10, 11
SB_WARPS, SB_HDR_INTS = 16, 16


def plan_array(graph_p, arr, n_head=0):
    L = swgn.lib()
    L.swgn_plan_array.argtypes = [C.POINTER(swgn.Graph), C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    n = C.c_int64()
    assert L.swgn_plan_array(graph_p, n_head, arr, None, C.byref(n)) == 0
    out = np.zeros(max(n.value, 1), np.int32)
    assert L.swgn_plan_array(graph_p, n_head, arr, out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)) == 0
    return out[:n.value]


def array_ids():
    """IArr ids of the three stream arrays, read from the header so that the test follows the enum."""
    import os
    import re
    src = open(os.path.join(swgn.HERE, "csrc", "device_types.h")).read()
    body = src[src.index("enum IArr {"):src.index("NUM_IARR")]
    names = re.findall(r"^\s*(I_[A-Z0-9_]+)", body, re.M)
    return {n: i for i, n in enumerate(names)}


def stream_info(graph_p, n_head=0):
    L = swgn.lib()
    L.swgn_plan_stream_info.argtypes = [C.POINTER(swgn.Graph), C.c_int32, C.POINTER(C.c_int32)]
    info = (C.c_int32 * 16)()
    assert L.swgn_plan_stream_info(graph_p, n_head, info) == 0
    keys = ["ok", "nbatch", "acc", "jcap", "rcap", "ecap", "fcap", "seccap", "n_fb", "smem"]
    return {k: info[i] for i, k in enumerate(keys)}


def interpret(graph_p, J, r, D, n_head=0):
    """Returns (S upper + rhs as the kernel would write them, E-buffer, chunk factors) or None when the window does
    not fit the on-chip budget."""
    ids = array_ids()
    st, pd = swgn.plan_probe(graph_p, n_head)
    assert st == 0
    info = stream_info(graph_p, n_head)
    if not info["ok"]:
        return None
    A = {k: plan_array(graph_p, ids[k], n_head) for k in
         ["I_COL_SIZE", "I_COL_POS", "I_TCOL", "I_ROW_RES", "I_ROW_NRES", "I_ROW_CELL", "I_CELL_COL", "I_CELL_VAL",
          "I_SB_HDR", "I_SB_REC", "I_ACC_MAP"]}
    n_e, n_f, n_ecols, n_cols = pd["n_e"], pd["n_f"], pd["n_ecols"], pd["n_cols"]
    n_fb = info["n_fb"]
    # W_JAC / W_RES as k_eval lays them out: cells in row order, row-major n_res x col_size
    JAC = np.zeros(pd["n_jac"] + 8)
    for row in range(pd["n_rows"]):
        r0, nr = A["I_ROW_RES"][row], A["I_ROW_NRES"][row]
        for c in range(A["I_ROW_CELL"][row], A["I_ROW_CELL"][row + 1]):
            col = A["I_CELL_COL"][c]
            pos, sz = A["I_COL_POS"][col], A["I_COL_SIZE"][col]
            JAC[A["I_CELL_VAL"][c]:A["I_CELL_VAL"][c] + nr * sz] = J[r0:r0 + nr, pos:pos + sz].ravel()
    RES = np.concatenate([r, np.zeros(8)])
    jcap, rcap, ecap, fcap = info["jcap"], info["rcap"], info["ecap"], info["fcap"]
    scap = jcap + rcap
    OA = np.full(2 * scap + ecap + fcap, np.nan)
    ACC = np.full(info["acc"], np.nan)
    EBUF, EFAC = {}, {}
    hdr = A["I_SB_HDR"].reshape(-1, SB_HDR_INTS)
    assert len(hdr) == info["nbatch"]
    REC = A["I_SB_REC"]
    n_terms = 0

    def runs(pkg, p0, p1):
        nonlocal n_terms
        u = p0
        while u < p1:
            dst, dst2, z, meta = (int(x) for x in pkg[u:u + 4])
            u += 4
            n_pos, n_neg, first, ecell = z & 0xfff, (z >> 12) & 0xfff, (z >> 24) & 1, (z >> 25) & 1
            n = n_pos + n_neg
            ps, qs, ti, tj, diag = meta & 63, (meta >> 6) & 63, ((meta >> 12) & 7) * 8, ((meta >> 15) & 7) * 8, (meta >> 18) & 1
            assert n_pos % 4 == 0 and n_neg % 4 == 0
            tile = np.zeros((8, 8))
            for t in range(n):
                w0, w1 = int(pkg[u]) & 0xffffffff, int(pkg[u + 1]) & 0xffffffff
                u += 2
                mask = (w1 >> 16) & 15
                if mask == 0:
                    assert w0 == 0 and w1 == 0
                    continue
                n_terms += 1
                a, bo, b2 = w0 & 0xffff, w0 >> 16, w1 & 0xffff
                m, sign = {1: 1, 3: 2, 7: 3, 15: 4}[mask], 1.0 if t < n_pos else -1.0
                Am = OA[a:a + m * ps].reshape(m, ps)
                Bm = OA[bo:bo + m * qs].reshape(m, qs)
                if diag:
                    Bm = np.hstack([Bm, OA[b2:b2 + m].reshape(m, 1)])
                full = sign * Am.T @ Bm
                assert np.all(np.isfinite(full)), "term reads an operand that has not been produced"
                blk = full[ti:ti + 8, tj:tj + 8]
                tile[:blk.shape[0], :blk.shape[1]] += blk
            stride = qs + diag
            for i in range(ti, min(ti + 8, ps)):
                for j in range(tj, min(tj + 8, stride)):
                    if ecell:
                        addr = dst + i * qs + j if j < qs else dst2 + i
                        OA[addr] = tile[i - ti, j - tj]
                    else:
                        addr = dst + i * stride + j
                        if first:
                            ACC[addr] = tile[i - ti, j - tj]
                        else:
                            assert np.isfinite(ACC[addr]), "accumulator read before its first run"
                            ACC[addr] += tile[i - ti, j - tj]

    for k, h in enumerate(hdr):
        rec_off, rec_len, j_src, j_len, r_src, r_len, eb_src, eb_len, ef_src, ef_len = (int(x) for x in h[:10])
        n_tchunk, off_textra, n_mchunk = int(h[10]), int(h[11]), int(h[12])
        off_tchunk, off_mchunk, sec_len = int(h[13]), int(h[14]), int(h[15])
        assert rec_off % 4 == 0 and rec_len % 4 == 0 and sec_len % 4 == 0 and j_src % 2 == 0 and r_src % 2 == 0
        assert j_len <= jcap and r_len <= rcap and eb_len <= ecap and ef_len <= fcap and sec_len <= info["seccap"]
        pkg = REC[rec_off:rec_off + rec_len]
        s = (k & 1) * scap
        OA[s:s + scap] = np.nan
        OA[s:s + j_len] = JAC[j_src:j_src + j_len]
        OA[s + jcap:s + jcap + r_len] = RES[r_src:r_src + r_len]
        wb, fb = 2 * scap, 2 * scap + ecap
        OA[wb:] = np.nan
        # A: landmark-like chunks (factor, w_g, W row by row)
        for c in range(n_tchunk):
            trow, y, z, epos = (int(x) for x in pkg[off_tchunk + 4 * c:off_tchunk + 4 * c + 4])
            nrows, es = y & 0xffff, y >> 16
            ete = np.diag(D[epos:epos + es] ** 2)
            g = np.zeros(es)
            rows = []
            for rr in range(nrows):
                x0, res, fw, w3 = (int(x) & 0xffffffff for x in pkg[trow + 4 * rr:trow + 4 * rr + 4])
                eo, nres = x0 & 0xffff, x0 >> 16
                E = OA[eo:eo + nres * es].reshape(nres, es)
                ete += E.T @ E
                g += E.T @ OA[res:res + nres]
                rows.append((E, nres, fw & 0xffff, fw >> 16, w3 & 0xff, (w3 >> 8) & 0xff, w3 >> 16))
            Lm = np.linalg.cholesky(ete)
            fo, go = z & 0xffff, (z >> 16) & 0xffff
            OA[fo:fo + es * es] = Lm.ravel()
            OA[go:go + es] = np.linalg.solve(Lm, g)
            for E, nres, f_oa, w_oa, fs, nfc, xi in rows:
                V = np.linalg.solve(Lm, E.T)
                for q in range(nfc):
                    fc = (f_oa, w_oa, fs) if q == 0 else tuple(int(x) for x in pkg[off_textra + 4 * (xi + q - 1):off_textra + 4 * (xi + q - 1) + 3])
                    F = OA[fc[0]:fc[0] + nres * fc[2]].reshape(nres, fc[2])
                    OA[fc[1]:fc[1] + es * fc[2]] = (V @ F).ravel()
        # A: raw products of the other chunks
        for wv in range(SB_WARPS):
            runs(pkg, int(pkg[wv]), int(pkg[wv + 1]))
        # B: other chunks
        for c in range(n_mchunk):
            es, epos, fo, ns1 = (int(x) for x in pkg[off_mchunk + 8 * c:off_mchunk + 8 * c + 4])
            slot0 = int(pkg[off_mchunk + 8 * c + 4])
            raw = OA[fo:fo + es * es].reshape(es, es)
            ete = np.triu(raw) + np.triu(raw, 1).T + np.diag(D[epos:epos + es] ** 2)
            Lm = np.linalg.cholesky(ete)
            OA[fo:fo + es * es] = Lm.ravel()
            for sidx in range(ns1):
                off, fs = int(pkg[slot0 + 2 * sidx]), int(pkg[slot0 + 2 * sidx + 1])
                Bm = OA[off:off + es * fs].reshape(es, fs)
                OA[off:off + es * fs] = np.linalg.solve(Lm, Bm).ravel()
        for x in range(eb_len):
            EBUF[eb_src + x] = OA[wb + x]
        for x in range(ef_len):
            EFAC[ef_src + x] = OA[fb + x]
        # C
        for wv in range(SB_WARPS):
            runs(pkg, int(pkg[SB_WARPS + 1 + wv]), int(pkg[SB_WARPS + 2 + wv]))
    # write-out
    S = np.zeros((n_f, n_f))
    rhs = np.zeros(n_f)
    amap = A["I_ACC_MAP"].reshape(n_fb, n_fb)
    for i in range(n_f):
        ci = A["I_TCOL"][n_e + i]
        p, li, ps = ci - n_ecols, n_e + i - A["I_COL_POS"][ci], A["I_COL_SIZE"][ci]
        rhs[i] = ACC[amap[p, p] + li * (ps + 1) + ps]
        for j in range(i, n_f):
            cj = A["I_TCOL"][n_e + j]
            q, lj, qs = cj - n_ecols, n_e + j - A["I_COL_POS"][cj], A["I_COL_SIZE"][cj]
            off = amap[p, q]
            if off >= 0:
                S[i, j] = ACC[off + li * (qs + (1 if p == q else 0)) + lj]
        S[i, i] += D[n_e + i] ** 2
    return S, rhs, n_terms, info


@pytest.mark.parametrize("which,wid,kw", [(1, 0, {}), (1, 3, {}),
                                          (2, 0, dict(n_keyframes=10, n_landmarks=100, n_gnss_epochs=5)),
                                          (2, 1, dict(n_keyframes=8, n_landmarks=40, n_gnss_epochs=8)),
                                          (2, 0, {})])
def test_stream_plan_reproduces_the_schur_complement(which, wid, kw):
    w = swgn.SynthWindow(which, wid, **kw)
    opt = w.options()
    o = ob.OracleSolver(w.graph_p, opt)
    _, r, _, J = o.evaluate()
    rng = np.random.default_rng(wid)
    D = rng.uniform(0.5, 1.5, o.n_cols) * 1e-2
    st, _, oS, orhs = o.linear_solve(D)
    assert st == 0
    out = interpret(w.graph_p, J, r, D)
    assert out is not None, "window does not fit the on-chip budget: %r" % (stream_info(w.graph_p),)
    S, rhs, n_terms, info = out
    assert np.all(np.isfinite(S)) and np.all(np.isfinite(rhs))
    scale = np.abs(oS).max()
    assert np.abs(np.triu(S) - np.triu(oS)).max() < 1e-12 * scale
    assert np.linalg.norm(rhs - orhs) < 1e-10 * max(np.linalg.norm(orhs), 1e-300)
    assert info["smem"] <= 227 * 1024


def test_windows_beyond_the_on_chip_budget_are_left_to_the_gather_kernel():
    w = swgn.SynthWindow(2, 0, n_keyframes=40, n_landmarks=300, n_gnss_epochs=20)
    info = stream_info(w.graph_p)
    assert info["ok"] == 0 and info["smem"] > 227 * 1024
