// Host planner of the streamed Schur elimination (device_types.h "streamed Schur", k_schur_stream.cu): cuts the
// window's rows into batches of whole chunks, lays out the on-chip operand area and the compact accumulators of
// the touched block cells of S, and writes the record packages the kernel interprets.  Integer work only.
//
// Same mathematics as the gather plan in plan.cpp (SchurEliminator<-1,-1,-1>::Eliminate,
// CERES/internal/ceres/schur_eliminator_impl.h:177-306): per chunk L L' = D_e^2 + sum E'E, W_f = L^-1 sum E'F_f,
// w_g = L^-1 sum E'b; S_pq = [p == q] D_p^2 + sum_rows F_p'F_q - sum_chunks W_p'W_q, rhs_p = sum F_p'b - sum W_p'w_g.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <numeric>

#include "plan.h"

namespace swgn {
namespace {

inline int al2(int x) { return (x + 1) & ~1; }
inline int al4(int x) { return (x + 3) & ~3; }

struct Unit {
  int r0, r1;   // rows
  int chunk;    // -1: a row without e-block
  int j0, j1;   // W_JAC range
  int s0, s1;   // W_RES range
  int eb0, eb1, ef0, ef1;  // W_EBUF / W_EFAC ranges (chunks)
  int n_terms;  // phase-C + phase-A terms (record budget)
  bool simple;  // class T: e-size <= 3, every slot fed by exactly one row, <= 2 residuals per row
};

struct Term {
  uint32_t a, b, b2, m, sign;  // OA offsets, rows (<= 4 after slab splitting)
};

struct Job {
  int dst, dst2, meta, first, ecell;
  const std::vector<Term>* terms;
};

}  // namespace

bool stream_enabled() {
  static const bool enabled = std::getenv("SWGN_SCHUR_STREAM") && std::atoi(std::getenv("SWGN_SCHUR_STREAM")) != 0;
  return enabled;
}

size_t stream_smem_bytes(int nbatch, int acc, int jcap, int rcap, int ecap, int fcap, int seccap) {
  // [2 mbarriers | per-warp record rings | batch headers | 2 section buffers | operand area | accumulators], 16-byte aligned parts
  size_t b = 16 + (size_t)SB_WARPS * SB_RING_BYTES;
  b += sizeof(int32_t) * (size_t)SB_HDR_INTS * (size_t)nbatch;
  b += sizeof(int32_t) * 2 * (size_t)al4(seccap);
  b += sizeof(double) * (size_t)(2 * (jcap + rcap) + ecap + fcap);
  b += sizeof(double) * (size_t)acc;
  return b;
}

void build_stream_plan(WindowPlan* P, int n_rows, int n_cols, int n_ecols, int n_jac, int n_res, const std::vector<int>& col_size,
                       const std::vector<int>& col_pos) {
  std::vector<int32_t>* I = P->iarr;
  int n_f_total = 0;
  for (int c = n_ecols; c < n_cols; ++c) n_f_total += col_size[c];
  StreamPlanInfo& sb = P->sb;
  sb = StreamPlanInfo();
  const int n_chunks = (int)I[I_CHUNK_ECOL].size();
  const int n_fb = n_cols - n_ecols;
  sb.n_fb = n_fb;
  if (n_fb <= 0 || n_fb > 2048) return;
  auto row_c0 = [&](int r) { return I[I_ROW_CELL][r]; };
  auto row_c1 = [&](int r) { return I[I_ROW_CELL][r + 1]; };
  auto row_j0 = [&](int r) { return r < n_rows ? I[I_CELL_VAL][row_c0(r)] : n_jac; };
  auto row_s0 = [&](int r) { return r < n_rows ? I[I_ROW_RES][r] : n_res; };

  // ---- units
  std::vector<Unit> units;
  {
    int r = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
      Unit u;
      u.r0 = I[I_CHUNK_ROW][ch];
      u.r1 = I[I_CHUNK_ROW][ch + 1];
      u.chunk = ch;
      const int es = col_size[I[I_CHUNK_ECOL][ch]];
      const int sl0 = I[I_CHUNK_SLOT][ch], sl1 = I[I_CHUNK_SLOT][ch + 1];
      u.eb0 = sl1 > sl0 ? I[I_SLOT_BUF][sl0] : I[I_CHUNK_G][ch];
      u.eb1 = al2(I[I_CHUNK_G][ch] + es);
      u.ef0 = I[I_CHUNK_FAC][ch];
      u.ef1 = u.ef0 + al2(es * es);
      u.simple = I[I_CHUNK_SIMPLE][ch] != 0 && es <= 3;
      for (int rr = u.r0; rr < u.r1 && u.simple; ++rr)
        if (I[I_ROW_NRES][rr] > 2) u.simple = false;
      units.push_back(u);
      r = u.r1;
    }
    for (; r < n_rows; ++r) {
      Unit u;
      u.r0 = r;
      u.r1 = r + 1;
      u.chunk = -1;
      u.eb0 = u.eb1 = u.ef0 = u.ef1 = 0;
      u.simple = false;
      units.push_back(u);
    }
    for (Unit& u : units) {
      u.j0 = row_j0(u.r0);
      u.j1 = row_j0(u.r1);
      u.s0 = row_s0(u.r0);
      u.s1 = row_s0(u.r1);
      int nt = 0;
      for (int rr = u.r0; rr < u.r1; ++rr) {
        const int slabs = (I[I_ROW_NRES][rr] + 3) / 4;
        int nfc = 0;
        for (int c = row_c0(rr); c < row_c1(rr); ++c) nfc += I[I_CELL_COL][c] >= n_ecols;
        nt += slabs * nfc * (nfc + 1) / 2;
        if (u.chunk >= 0 && !u.simple) nt += slabs * (1 + nfc);
      }
      if (u.chunk >= 0) {
        const int es = col_size[I[I_CHUNK_ECOL][u.chunk]];
        const int ns = I[I_CHUNK_SLOT][u.chunk + 1] - I[I_CHUNK_SLOT][u.chunk];
        nt += ((es + 3) / 4) * ns * (ns + 1) / 2;
      }
      u.n_terms = nt;
    }
  }

  // ---- batches: greedy packing of consecutive units
  struct Batch {
    int u0, u1;
    int j_src, j_len, r_src, r_len, eb_src, eb_len, ef_src, ef_len;
    int part, nparts;  // a single row with more terms than a package holds is cut into nparts batches (same J segment,
                       // disjoint sets of block cells)
  };
  std::vector<Batch> batches;
  {
    size_t k = 0;
    while (k < units.size()) {
      Batch bt;
      bt.u0 = (int)k;
      const int jb = units[k].j0 & ~1, sb0 = units[k].s0 & ~1;
      int eb0 = -1, eb1 = 0, ef0 = -1, ef1 = 0, terms = 0;
      size_t e = k;
      while (e < units.size()) {
        const Unit& u = units[e];
        const int jl = al2(u.j1 - jb);
        const int t2 = terms + u.n_terms;
        if (e > k && (jl > SB_JCAP || t2 > SB_TERMCAP)) break;
        terms = t2;
        if (u.chunk >= 0) {
          if (eb0 < 0) { eb0 = u.eb0; ef0 = u.ef0; }
          eb1 = u.eb1;
          ef1 = u.ef1;
        }
        ++e;
      }
      bt.u1 = (int)e;
      bt.j_src = jb;
      bt.j_len = al2(units[e - 1].j1 - jb);
      bt.r_src = sb0;
      bt.r_len = al2(units[e - 1].s1 - sb0);
      bt.eb_src = eb0 < 0 ? 0 : eb0;
      bt.eb_len = eb0 < 0 ? 0 : eb1 - eb0;
      bt.ef_src = ef0 < 0 ? 0 : ef0;
      bt.ef_len = ef0 < 0 ? 0 : ef1 - ef0;
      bt.part = 0;
      bt.nparts = 1;
      if (e == k + 1 && units[k].chunk < 0 && units[k].n_terms > SB_TERMCAP) bt.nparts = (units[k].n_terms + SB_TERMCAP - 1) / SB_TERMCAP;
      for (int part = 0; part < bt.nparts; ++part) {
        bt.part = part;
        batches.push_back(bt);
      }
      k = e;
    }
  }
  int jcap = 2, rcap = 2, ecap = 2, fcap = 2;
  for (const Batch& bt : batches) {
    jcap = std::max(jcap, bt.j_len);
    rcap = std::max(rcap, bt.r_len);
    ecap = std::max(ecap, bt.eb_len);
    fcap = std::max(fcap, bt.ef_len);
  }
  const int scap = jcap + rcap;
  if (2 * scap + ecap + fcap > 65535) return;  // 16-bit operand offsets

  // ---- compact accumulators of the touched block cells (p <= q) of S
  std::vector<int32_t>& amap = I[I_ACC_MAP];
  amap.assign((size_t)n_fb * n_fb, -1);
  auto cid = [&](int p, int q) { return (size_t)(p - n_ecols) * n_fb + (size_t)(q - n_ecols); };
  {
    std::vector<char> touched((size_t)n_fb * n_fb, 0);
    for (int r = 0; r < n_rows; ++r)
      for (int c1 = row_c0(r); c1 < row_c1(r); ++c1) {
        const int p = I[I_CELL_COL][c1];
        if (p < n_ecols) continue;
        for (int c2 = c1; c2 < row_c1(r); ++c2) touched[cid(p, I[I_CELL_COL][c2])] = 1;
      }
    for (int ch = 0; ch < n_chunks; ++ch)
      for (int s1 = I[I_CHUNK_SLOT][ch]; s1 < I[I_CHUNK_SLOT][ch + 1]; ++s1)
        for (int s2 = s1; s2 < I[I_CHUNK_SLOT][ch + 1]; ++s2) touched[cid(I[I_SLOT_COL][s1], I[I_SLOT_COL][s2])] = 1;
    int off = 0;
    for (int p = n_ecols; p < n_cols; ++p)
      for (int q = p; q < n_cols; ++q) {
        if (!touched[cid(p, q)] && p != q) continue;
        if (col_size[p] > MAX_COL_SIZE || col_size[q] > MAX_COL_SIZE) return;
        amap[cid(p, q)] = off;
        off += col_size[p] * (col_size[q] + (p == q ? 1 : 0));
      }
    sb.acc = std::max(256, al2(off));  // (unpredicated operand loads may run up to ~200 doubles past the operand area)
  }
  // first touch of every accumulator tile: the run that sees it first starts from zero instead of reading it
  std::vector<int> tile_base((size_t)n_fb * n_fb, -1);
  int n_tiles = 0;
  for (int p = n_ecols; p < n_cols; ++p)
    for (int q = p; q < n_cols; ++q)
      if (amap[cid(p, q)] >= 0) {
        tile_base[cid(p, q)] = n_tiles;
        n_tiles += ((col_size[p] + 7) / 8) * ((col_size[q] + (p == q ? 1 : 0) + 7) / 8);
      }
  std::vector<char> tile_seen(n_tiles, 0);

  // ---- record packages
  std::vector<int32_t>& HDR = I[I_SB_HDR];
  std::vector<int32_t>& REC = I[I_SB_REC];
  HDR.clear();
  REC.clear();
  int reccap = 0;
  std::vector<std::pair<size_t, Term>> cterms;  // (cell id, term) of the current batch
  std::deque<std::vector<Term>> term_lists;  // stable addresses
  std::vector<Job> jobs_a, jobs_c;
  std::vector<int32_t> pk;
  for (size_t bi = 0; bi < batches.size(); ++bi) {
    const Batch& bt = batches[bi];
    const int stage = (int)(bi & 1) * scap;
    const int wb = 2 * scap, fb = 2 * scap + ecap;
    auto oaJ = [&](int j) { return (uint32_t)(stage + (j - bt.j_src)); };
    auto oaR = [&](int s) { return (uint32_t)(stage + jcap + (s - bt.r_src)); };
    auto oaW = [&](int e) { return (uint32_t)(wb + (e - bt.eb_src)); };
    auto oaF = [&](int f) { return (uint32_t)(fb + (f - bt.ef_src)); };
    cterms.clear();
    term_lists.clear();
    jobs_a.clear();
    jobs_c.clear();
    std::vector<int32_t> sec_tchunk, sec_trow, sec_textra, sec_mchunk, sec_slot;
    bool too_many = false;
    for (int ui = bt.u0; ui < bt.u1; ++ui) {
      const Unit& u = units[ui];
      // F'F terms of the rows
      for (int r = u.r0; r < u.r1; ++r) {
        const uint32_t nres = (uint32_t)I[I_ROW_NRES][r];
        for (int c1 = row_c0(r); c1 < row_c1(r); ++c1) {
          const int p = I[I_CELL_COL][c1];
          if (p < n_ecols) continue;
          for (int c2 = c1; c2 < row_c1(r); ++c2) {
            const int q = I[I_CELL_COL][c2];
            for (uint32_t e0 = 0; e0 < nres; e0 += 4)
              cterms.push_back({cid(p, q), Term{oaJ(I[I_CELL_VAL][c1]) + e0 * (uint32_t)col_size[p], oaJ(I[I_CELL_VAL][c2]) + e0 * (uint32_t)col_size[q],
                                                oaR(I[I_ROW_RES][r]) + e0, std::min(4u, nres - e0), 0u}});
          }
        }
      }
      if (u.chunk < 0) continue;
      const int ch = u.chunk;
      const int ecol = I[I_CHUNK_ECOL][ch];
      const int es = col_size[ecol];
      const int sl0 = I[I_CHUNK_SLOT][ch], sl1 = I[I_CHUNK_SLOT][ch + 1];
      // -W'W terms of the chunk
      for (int s1 = sl0; s1 < sl1; ++s1)
        for (int s2 = s1; s2 < sl1; ++s2) {
          const int p = I[I_SLOT_COL][s1], q = I[I_SLOT_COL][s2];
          for (uint32_t e0 = 0; e0 < (uint32_t)es; e0 += 4)
            cterms.push_back({cid(p, q), Term{oaW(I[I_SLOT_BUF][s1]) + e0 * (uint32_t)col_size[p], oaW(I[I_SLOT_BUF][s2]) + e0 * (uint32_t)col_size[q],
                                              oaW(I[I_CHUNK_G][ch]) + e0, std::min(4u, (uint32_t)es - e0), 1u}});
        }
      if (u.simple) {
        // class T: a group of 8 lanes per chunk -- partial E'E / E'b per lane, shuffle reduction, L and w_g, then W row by row
        const int32_t rec[4] = {(int32_t)(sec_trow.size() / 4), (u.r1 - u.r0) | (es << 16), (int32_t)(oaF(I[I_CHUNK_FAC][ch]) | (oaW(I[I_CHUNK_G][ch]) << 16)),
                                col_pos[ecol]};
        sec_tchunk.insert(sec_tchunk.end(), rec, rec + 4);
        for (int r = u.r0; r < u.r1; ++r) {
          const int c0 = row_c0(r), nfc = row_c1(r) - c0 - 1, nres = I[I_ROW_NRES][r];
          const uint32_t f_oa = nfc > 0 ? oaJ(I[I_CELL_VAL][c0 + 1]) : 0u, w_oa = nfc > 0 ? oaW(I[I_CELL_SLOT][c0 + 1]) : 0u;
          const int fs = nfc > 0 ? col_size[I[I_CELL_COL][c0 + 1]] : 0;
          if (nfc > 255 || sec_textra.size() / 4 > 0xffff) too_many = true;
          const int32_t rr[4] = {(int32_t)(oaJ(I[I_CELL_VAL][c0]) | ((uint32_t)nres << 16)), (int32_t)oaR(I[I_ROW_RES][r]), (int32_t)(f_oa | (w_oa << 16)),
                                 (int32_t)((uint32_t)fs | ((uint32_t)std::max(nfc, 0) << 8) | ((uint32_t)(sec_textra.size() / 4) << 16))};
          sec_trow.insert(sec_trow.end(), rr, rr + 4);
          for (int c = c0 + 2; c < row_c1(r); ++c) {
            const int32_t x[4] = {(int32_t)oaJ(I[I_CELL_VAL][c]), (int32_t)oaW(I[I_CELL_SLOT][c]), col_size[I[I_CELL_COL][c]], 0};
            sec_textra.insert(sec_textra.end(), x, x + 4);
          }
        }
      } else {
        // class M: raw products E'[E | b | F] on the tensor pipe (phase A), factor + forward substitution by one warp (phase B)
        const int32_t m0[4] = {es, col_pos[ecol], (int32_t)oaF(I[I_CHUNK_FAC][ch]), sl1 - sl0 + 1};
        const int32_t m1[4] = {(int32_t)(sec_slot.size() / 2), 0, 0, 0};
        sec_mchunk.insert(sec_mchunk.end(), m0, m0 + 4);
        sec_mchunk.insert(sec_mchunk.end(), m1, m1 + 4);
        for (int s = sl0; s < sl1; ++s) {
          sec_slot.push_back((int32_t)oaW(I[I_SLOT_BUF][s]));
          sec_slot.push_back(col_size[I[I_SLOT_COL][s]]);
        }
        sec_slot.push_back((int32_t)oaW(I[I_CHUNK_G][ch]));
        sec_slot.push_back(1);
        // [E'E | E'b]
        term_lists.emplace_back();
        std::vector<Term>& dt = term_lists.back();
        for (int r = u.r0; r < u.r1; ++r) {
          const uint32_t nres = (uint32_t)I[I_ROW_NRES][r], eo = oaJ(I[I_CELL_VAL][row_c0(r)]);
          for (uint32_t e0 = 0; e0 < nres; e0 += 4)
            dt.push_back(Term{eo + e0 * (uint32_t)es, eo + e0 * (uint32_t)es, oaR(I[I_ROW_RES][r]) + e0, std::min(4u, nres - e0), 0u});
        }
        for (int ti = 0; ti < es; ti += 8)
          for (int tj = 0; tj < es + 1; tj += 8) {
            if (tj + 7 < ti) continue;
            jobs_a.push_back(Job{(int)oaF(I[I_CHUNK_FAC][ch]), (int)oaW(I[I_CHUNK_G][ch]), es | (es << 6) | ((ti / 8) << 12) | ((tj / 8) << 15) | (1 << 18), 1, 1, &dt});
          }
        // E'F per slot
        for (int s = sl0; s < sl1; ++s) {
          term_lists.emplace_back();
          std::vector<Term>& st = term_lists.back();
          const int fs = col_size[I[I_SLOT_COL][s]];
          for (int r = u.r0; r < u.r1; ++r) {
            const uint32_t nres = (uint32_t)I[I_ROW_NRES][r], eo = oaJ(I[I_CELL_VAL][row_c0(r)]);
            for (int c = row_c0(r) + 1; c < row_c1(r); ++c)
              if (I[I_CELL_SLOT][c] == I[I_SLOT_BUF][s])
                for (uint32_t e0 = 0; e0 < nres; e0 += 4)
                  st.push_back(Term{eo + e0 * (uint32_t)es, oaJ(I[I_CELL_VAL][c]) + e0 * (uint32_t)fs, 0u, std::min(4u, nres - e0), 0u});
          }
          for (int ti = 0; ti < es; ti += 8)
            for (int tj = 0; tj < fs; tj += 8)
              jobs_a.push_back(Job{(int)oaW(I[I_SLOT_BUF][s]), 0, es | (fs << 6) | ((ti / 8) << 12) | ((tj / 8) << 15), 1, 1, &st});
        }
      }
    }
    // phase-C jobs: the batch's terms grouped by block cell, one job per 8x8 tile of the cell
    std::stable_sort(cterms.begin(), cterms.end(), [](const std::pair<size_t, Term>& x, const std::pair<size_t, Term>& y) { return x.first < y.first; });
    static const std::vector<Term> kNoTerms;
    const size_t per_part = (cterms.size() + (size_t)bt.nparts - 1) / (size_t)bt.nparts;
    for (size_t k = 0; k < cterms.size();) {
      size_t e = k;
      while (e < cterms.size() && cterms[e].first == cterms[k].first) ++e;
      if (bt.nparts > 1 && (int)std::min<size_t>(k / std::max<size_t>(per_part, 1), (size_t)bt.nparts - 1) != bt.part) {
        k = e;
        continue;
      }
      term_lists.emplace_back();
      std::vector<Term>& ts = term_lists.back();
      ts.reserve(e - k);
      for (size_t x = k; x < e; ++x) ts.push_back(cterms[x].second);
      const size_t c = cterms[k].first;
      const int p = n_ecols + (int)(c / n_fb), q = n_ecols + (int)(c % n_fb);
      const int ps = col_size[p], qs = col_size[q], diag = p == q ? 1 : 0;
      const int ntj = (qs + diag + 7) / 8;
      for (int ti = 0; ti < ps; ti += 8)
        for (int tj = 0; tj < qs + diag; tj += 8) {
          if (diag && tj + 7 < ti) continue;  // entirely below the diagonal of a diagonal cell: never read
          const int tile = tile_base[c] + (ti / 8) * ntj + tj / 8;
          jobs_c.push_back(Job{amap[c], 0, ps | (qs << 6) | ((ti / 8) << 12) | ((tj / 8) << 15) | (diag << 18), tile_seen[tile] ? 0 : 1, 0, &ts});
          tile_seen[tile] = 1;
        }
      k = e;
    }
    if (bi + 1 == batches.size()) {
      // accumulator tiles no term ever touches (diagonal cells that only carry D^2) still have to read as zero
      for (int p = n_ecols; p < n_cols; ++p)
        for (int q = p; q < n_cols; ++q) {
          const size_t c = cid(p, q);
          if (amap[c] < 0) continue;
          const int ps = col_size[p], qs = col_size[q], diag = p == q ? 1 : 0, ntj = (qs + diag + 7) / 8;
          for (int ti = 0; ti < ps; ti += 8)
            for (int tj = 0; tj < qs + diag; tj += 8) {
              if (diag && tj + 7 < ti) continue;
              const int tile = tile_base[c] + (ti / 8) * ntj + tj / 8;
              if (tile_seen[tile]) continue;
              jobs_c.push_back(Job{amap[c], 0, ps | (qs << 6) | ((ti / 8) << 12) | ((tj / 8) << 15) | (diag << 18), 1, 0, &kNoTerms});
              tile_seen[tile] = 1;
            }
        }
    }
    // ---- assemble the package
    const int n_ptr = 2 * (SB_WARPS + 1);
    pk.assign((size_t)al4(n_ptr), 0);
    auto append = [&](const std::vector<int32_t>& sec) {
      const int off = (int)pk.size();
      pk.insert(pk.end(), sec.begin(), sec.end());
      pk.resize((size_t)al4((int)pk.size()), 0);
      return off;
    };
    const int off_tchunk = append(sec_tchunk);
    const int off_trow = append(sec_trow);
    const int off_textra = append(sec_textra);
    const int off_mchunk = append(sec_mchunk);
    const int off_slot = append(sec_slot);
    // absolute int offsets of the nested lists
    for (size_t k = 0; k < sec_tchunk.size() / 4; ++k) pk[off_tchunk + 4 * k] = off_trow + 4 * pk[off_tchunk + 4 * k];
    for (size_t k = 0; k < sec_mchunk.size() / 8; ++k) pk[off_mchunk + 8 * k + 4] = off_slot + 2 * pk[off_mchunk + 8 * k + 4];
    bool too_long = false;
    auto deal = [&](const std::vector<Job>& jobs, int ptr0) {
      std::vector<size_t> idx(jobs.size());
      std::iota(idx.begin(), idx.end(), 0);
      std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return jobs[x].terms->size() > jobs[y].terms->size(); });
      std::vector<std::vector<size_t>> mine(SB_WARPS);
      std::vector<size_t> load(SB_WARPS, 0);
      for (size_t j : idx) {
        const int wmin = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        mine[wmin].push_back(j);
        load[wmin] += (jobs[j].terms->size() + 3) / 4 * 4 + 8;
      }
      for (int wv = 0; wv < SB_WARPS; ++wv) {
        pk[ptr0 + wv] = (int32_t)pk.size();
        for (size_t j : mine[wv]) {
          const Job& jb = jobs[j];
          // + terms first, then - terms, each list padded to a multiple of 4 (a group of MMAs) with all-zero records
          size_t n_sign[2] = {0, 0};
          for (const Term& t : *jb.terms) ++n_sign[t.sign];
          const size_t np[2] = {(n_sign[0] + 3) / 4 * 4, (n_sign[1] + 3) / 4 * 4};
          if (np[0] > 0xfff || np[1] > 0xfff) too_long = true;
          pk.push_back(jb.dst);
          pk.push_back(jb.dst2);
          pk.push_back((int32_t)(np[0] | (np[1] << 12) | ((size_t)jb.first << 24) | ((size_t)jb.ecell << 25)));
          pk.push_back(jb.meta);
          for (uint32_t sg = 0; sg < 2; ++sg) {
            size_t e = 0;
            for (const Term& t : *jb.terms) {
              if (t.sign != sg) continue;
              pk.push_back((int32_t)(t.a | (t.b << 16)));
              pk.push_back((int32_t)(t.b2 | (((1u << t.m) - 1u) << 16)));  // row mask: rows beyond the slab read as zero
              ++e;
            }
            for (; e < np[sg]; ++e) {
              pk.push_back(0);
              pk.push_back(0);
            }
          }
        }
      }
      pk[ptr0 + SB_WARPS] = (int32_t)pk.size();
    };
    const int sec_len = (int)pk.size();  // what precedes goes to shared memory with the batch; the run streams are read from L2
    deal(jobs_a, 0);
    deal(jobs_c, SB_WARPS + 1);
    if (too_long) return;  // run length fields
    const int32_t hdr[SB_HDR_INTS] = {(int32_t)REC.size(), (int32_t)pk.size(), bt.j_src, bt.j_len, bt.r_src, bt.r_len, bt.eb_src, bt.eb_len,
                                      bt.ef_src, bt.ef_len, (int32_t)(sec_tchunk.size() / 4), off_textra,
                                      (int32_t)(sec_mchunk.size() / 8), off_tchunk, off_mchunk, sec_len};
    if (too_many) return;
    HDR.insert(HDR.end(), hdr, hdr + SB_HDR_INTS);
    REC.insert(REC.end(), pk.begin(), pk.end());
    reccap = std::max(reccap, sec_len);
  }
  sb.nbatch = (int)batches.size();
  sb.jcap = jcap;
  sb.rcap = rcap;
  sb.ecap = ecap;
  sb.fcap = fcap;
  sb.reccap = reccap;
  const size_t smem = stream_smem_bytes(sb.nbatch, sb.acc, sb.jcap, sb.rcap, sb.ecap, sb.fcap, sb.reccap);
  // the write-out of S looks the accumulators up through tables that reuse the section buffers and the operand area
  const size_t tables = sizeof(int32_t) * ((size_t)n_fb * n_fb + (size_t)al4(n_f_total) + 2 * (size_t)al4(n_fb));
  const size_t scratch = sizeof(int32_t) * 2 * (size_t)al4(reccap) + sizeof(double) * (size_t)(2 * (jcap + rcap) + ecap + fcap);
  // opt-in (SWGN_SCHUR_STREAM=1): on the B200 the streamed kernel cuts the DRAM traffic of the elimination to the
  // algorithmic bytes but is instruction / latency bound and, as measured, still slower than the gather kernel
  const bool enabled = stream_enabled();
  sb.fits = sb.nbatch > 0 && sb.nbatch <= 1024 && smem <= (size_t)SB_SMEM_BUDGET && tables <= scratch;
  sb.ok = enabled && sb.nbatch > 0 && sb.nbatch <= 1024 && smem <= (size_t)SB_SMEM_BUDGET && tables <= scratch;
}

}  // namespace swgn
