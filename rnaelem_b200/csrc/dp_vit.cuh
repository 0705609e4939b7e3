// dp_vit.cuh -- constrained Viterbi + traceback of the scanner (RNAelemScanDP::calc_viterbi_alignment,
// motif_scanner.hpp:172-184; CYKFun :802-913; trace_back :262-362) for sequences whose motif region [Ys,Ye] is known.
//
// The reference fills a max-plus table and, next to every entry, a 20-byte Trace{k,l,t,e1,s1_id} of the first strict
// maximum in visiting order, then walks the trace entries from the root.  Here
//   * the forward pass keeps VALUES only (max is exact and order independent, so the table is bit-identical to the
//     reference's whatever the evaluation order) -- no trace table, half the HBM traffic of the pass;
//   * the traceback re-enumerates the one (cell, state type, state) it pops, in the reference's visiting order
//     (dp_enum.cuh), and takes the first candidate that reaches the stored maximum: the same arg-max, ties included;
//   * states that cannot be part of a parse with the motif at [Ys,Ye] (StartEndConstraint::outside_cells) are neither
//     computed nor stored nor filled: a read of such an entry is answered with -inf by the index function (bidx);
//   * work mapping per diagonal: a cell whose bases all lie outside the motif region carries ONE state -> one thread per
//     such cell (32 cells per warp); a cell that overlaps the region carries S states -> one warp per cell, lane = state,
//     so the structural part of an enumeration (split points, inner pairs, loop energies) is evaluated once per warp in
//     lockstep instead of once per state.
// All sums use explicit IEEE adds in the reference's association (util.hpp:223-224), see MaxV in dp_pass.cuh.
#ifndef RELEM_DP_VIT_CUH
#define RELEM_DP_VIT_CUH
#include "dp_pass.cuh"
#include "dp_prim.cuh"

namespace relem {
namespace dp {

#define RELEM_VIT_NONE 0xFFFFFFFFu

// plane of the (first) child a transition type leads to; 7 = exterior row
RDEV int child_plane_of(int tt) {
  switch (tt) {
    case TT_E_H: return PL_L; case TT_P_E: return PL_E; case TT_P_P: return PL_P; case TT_O_O: return 7;
    case TT_O_OP: return PL_P; case TT_E_P: return PL_P; case TT_E_M: return PL_M; case TT_M_M: return PL_M;
    case TT_M_B: return PL_B; case TT_B_12: return PL_1; case TT_1_B: return PL_B; case TT_1_2: return PL_2;
    case TT_2_2: return PL_2; case TT_2_P: return PL_P; case TT_L_L: return PL_L;
  }
  return -1;
}

// which entries of the band exist for a fixed motif region
struct VitRegion {
  int ys, ye, s_bg0, s_bgM;
  bool on;
  RDEV void cells(int d, int ncell, int& nb, int& na) const {
    nb = 0; na = 0;
    if (!on || d == 0) return;
    nb = ys - d; nb = nb < 0 ? 0 : (nb > ncell ? ncell : nb);
    na = ncell - (ye + 2); na = na < 0 ? 0 : na;
    if (na > ncell - nb) na = ncell - nb;
  }
  // is entry (i, d, s) kept?  (same partition as cells(): leading i + d < ys, trailing i > ye + 1)
  RDEV bool kept(int i, int d, int s) const {
    if (!on || d == 0) return true;
    if (i + d < ys) return s == s_bg0;
    if (i > ye + 1) return s == s_bgM;
    return true;
  }
};

template <class CON> struct VitBase {
  const double* tab;
  const double* otab;
  CON con;
  VitRegion rg;
  RDEV bool allow(const ModelView& m, const SeqView& q, const Emit& e) const { return con.ok(m, q, e); }
  RDEV unsigned bidx(const SeqView& q, int plane, int i, int d, int s) const {
    return rg.kept(i, d, s) ? band_idx(q, plane, i, d, s) : RELEM_VIT_NONE;
  }
  RDEV double ld(unsigned c) const { return c == RELEM_VIT_NONE ? NINF : tab[c]; }
};

// forward: value of the maximum
template <class CON> struct VitV : VitBase<CON> {
  double best;
  RDEV void init() { best = NINF; }
  RDEV void cmp(double y) { if (best < y) best = y; }
  RDEV void t1(int, unsigned c0, double diff, double, int, const Emit&, const Geo&) { cmp(d_add(this->ld(c0), diff)); }
  RDEV void t2(int, unsigned c0, unsigned c1, double diff, double, int, const Emit&, const Geo&) {
    cmp(d_add(this->ld(c0), d_add(this->ld(c1), diff)));
  }
  RDEV void t3(int, unsigned c0, unsigned c1, unsigned c2, double diff, double, int, const Emit&, const Geo&) {
    cmp(d_add(this->ld(c0), d_add(this->ld(c1), d_add(this->ld(c2), diff))));
  }
  RDEV void o1(int, unsigned o0, double diff, double, int, const Emit&, const Geo&) { cmp(d_add(this->otab[o0], diff)); }
  RDEV void o2(int, unsigned o0, unsigned c1, double diff, double, int, const Emit&, const Geo&) {
    cmp(d_add(this->otab[o0], d_add(this->ld(c1), diff)));
  }
};

// traceback: first strict maximum in visiting order and its packed trace entry (pack_trace, dp_pass.cuh)
template <class CON> struct VitTraceV : VitBase<CON> {
  double best;
  unsigned long long tr;
  RDEV void init() { best = NINF; tr = RELEM_NO_TRACE; }
  RDEV void cmp(double y, int tt, const Geo& g) { if (best < y) { best = y; tr = pack_trace(tt, g); } }
  RDEV void t1(int tt, unsigned c0, double diff, double, int, const Emit&, const Geo& g) { cmp(d_add(this->ld(c0), diff), tt, g); }
  RDEV void t2(int tt, unsigned c0, unsigned c1, double diff, double, int, const Emit&, const Geo& g) {
    cmp(d_add(this->ld(c0), d_add(this->ld(c1), diff)), tt, g);
  }
  RDEV void t3(int tt, unsigned c0, unsigned c1, unsigned c2, double diff, double, int, const Emit&, const Geo& g) {
    cmp(d_add(this->ld(c0), d_add(this->ld(c1), d_add(this->ld(c2), diff))), tt, g);
  }
  RDEV void o1(int tt, unsigned o0, double diff, double, int, const Emit&, const Geo& g) { cmp(d_add(this->otab[o0], diff), tt, g); }
  RDEV void o2(int tt, unsigned o0, unsigned c1, double diff, double, int, const Emit&, const Geo& g) {
    cmp(d_add(this->otab[o0], d_add(this->ld(c1), diff)), tt, g);
  }
};

// the same-cell chain L, P, B, 2, 1, M, E of one (cell, state); same-cell reads (2 <- P, 1 <- 2 B, M <- B, E <- M L)
// are entries this thread has just written
template <class CON>
RDEV void vit_cell_state(const ModelView& m, const SeqView& q, double* tab, const double* otab, const CON& con,
                         const VitRegion& rg, int i, int d, int s) {
  const DevHMM& h = m.h;
  VitV<CON> v;
  v.tab = tab; v.otab = otab; v.con = con; v.rg = rg;
  {
    unsigned idx = band_idx(q, PL_L, i, d, s);
    if (d == 0) tab[idx] = (ld_ro(h.st_l + s) == ld_ro(h.st_r + s)) ? 0. : NINF;
    else { v.init(); enum_L(m, q, i, d, s, v); tab[idx] = v.best; }
  }
  if (d < q.min_pair - 2) return;   // no gate below opens before span 3 (E: d + 2 >= min_pair)
  const bool gP = ok_P(q, i, d), gB = ok_B(q, i, d), gM = ok_M(q, i, d), gE = ok_E(q, i, d);
  if (gP) { v.init(); enum_P(m, q, i, d, s, v); tab[band_idx(q, PL_P, i, d, s)] = v.best; }
  if (gB) {
    v.init(); enum_B(m, q, i, d, s, v); tab[band_idx(q, PL_B, i, d, s)] = v.best;
    v.init(); enum_2(m, q, i, d, s, v); tab[band_idx(q, PL_2, i, d, s)] = v.best;
    v.init(); enum_1(m, q, i, d, s, v); tab[band_idx(q, PL_1, i, d, s)] = v.best;
  }
  if (gM) { v.init(); enum_M(m, q, i, d, s, v); tab[band_idx(q, PL_M, i, d, s)] = v.best; }
  if (gE) { v.init(); enum_E(m, q, i, d, s, v); tab[band_idx(q, PL_E, i, d, s)] = v.best; }
}

// forward pass of one sequence by one CTA; one barrier per diagonal
template <class CON>
RDEV void cta_viterbi_forward(const ModelView& m, const SeqView& q, double* tab, double* otab, const CON& con,
                              const VitRegion& rg) {
  const int S = q.S, L = q.L, W = q.W;
  const int lane = lane_id(), w0 = warp_id(), nw = n_warps();
  for (int d = 0; d <= W; ++d) {
    const int ncell = L + 1 - d;
    int nb, na;
    rg.cells(d, ncell, nb, na);
    const int nbg = nb + na, npack = (nbg + WARP_N - 1) / WARP_N, nfull = ncell - nbg;
    // heavy units (full cells) first, the packs of single-state cells fill the tail
    for (int u = w0; u < nfull + npack; u += nw) {
      // one call site for both kinds of unit: (cell, first state, state stride, end)
      int i, s0, s1, ds;
      if (u < nfull) { i = nb + u; s0 = lane; s1 = S; ds = WARP_N; }
      else {
        const int t = (u - nfull) * WARP_N + lane;
        const bool lead = t < nb;
        i = lead ? t : ncell - na + (t - nb);
        s0 = lead ? rg.s_bg0 : rg.s_bgM;
        s1 = t < nbg ? s0 + 1 : s0;
        ds = 1;
      }
      for (int s = s0; s < s1; s += ds) vit_cell_state(m, q, tab, otab, con, rg, i, d, s);
    }
    CTA_SYNC();
  }
  // exterior recurrence (never pruned: S values per position)
  for (int t = CTA_TID; t < S; t += CTA_NTH) otab[t] = (t == m.h.s00) ? 0. : NINF;
  CTA_SYNC();
  for (int j = 1; j <= L; ++j) {
    for (int s = CTA_TID; s < S; s += CTA_NTH) {
      VitV<CON> v;
      v.tab = tab; v.otab = otab; v.con = con; v.rg = rg;
      v.init();
      enum_O(m, q, j, s, v);
      otab[j * S + s] = v.best;
    }
    CTA_SYNC();
  }
}

// RNAelemScanDP::trace_back (motif_scanner.hpp:262-362), one thread; the trace entry of the popped item is recomputed
template <class CON>
RDEV void vit_trace_back(const ModelView& m, const SeqView& q, const double* tab, const double* otab, const CON& con,
                         const VitRegion& rg, const int* n2s, int* stack, int s0, int* psihat, char* rss) {
  const DevHMM& h = m.h;
  const int M = h.M;
  int sp = 0;
#define PUSH(I, J, E, SS) { stack[sp * 4] = (I); stack[sp * 4 + 1] = (J); stack[sp * 4 + 2] = (E); stack[sp * 4 + 3] = (SS); ++sp; }
  PUSH(0, q.L, 7, s0)
  while (sp > 0) {
    --sp;
    const int ti = stack[sp * 4], tj = stack[sp * 4 + 1], te = stack[sp * 4 + 2], ts = stack[sp * 4 + 3];
    const int td = tj - ti;
    VitTraceV<CON> v;
    v.tab = tab; v.otab = otab; v.con = con; v.rg = rg;
    v.init();
    switch (te) {
      case 7: if (tj > 0) enum_O(m, q, tj, ts, v); break;
      case PL_L: if (td > 0) enum_L(m, q, ti, td, ts, v); break;
      case PL_P: enum_P(m, q, ti, td, ts, v); break;
      case PL_B: enum_B(m, q, ti, td, ts, v); break;
      case PL_2: enum_2(m, q, ti, td, ts, v); break;
      case PL_1: enum_1(m, q, ti, td, ts, v); break;
      case PL_M: enum_M(m, q, ti, td, ts, v); break;
      case PL_E: enum_E(m, q, ti, td, ts, v); break;
      default: break;
    }
    const unsigned long long tr = v.tr;
    if (tr == RELEM_NO_TRACE) continue;
    int tt = (int)(tr >> 56), s1 = (int)((tr >> 40) & 0xFFFF), k = (int)((tr >> 20) & 0xFFFFF), l = (int)(tr & 0xFFFFF);
    int e1 = child_plane_of(tt);
    int tsl = ld_ro(h.st_l + ts), tsr = ld_ro(h.st_r + ts);
    int s1l = ld_ro(h.st_l + s1), s1r = ld_ro(h.st_r + s1);
    switch (tt) {
      case TT_L_L: psihat[l] = tsr; PUSH(k, l, e1, s1) break;
      case TT_O_O: psihat[l] = tsr; rss[l] = 'O'; PUSH(k, l, e1, s1) break;
      case TT_2_2: psihat[l] = tsr; rss[l] = 'M'; PUSH(k, l, e1, s1) break;
      case TT_E_H: for (int p = ti; p < tj; ++p) rss[p] = 'H'; PUSH(k, l, e1, ts) break;
      case TT_E_M: case TT_M_B: case TT_2_P: case TT_1_2: case TT_1_B: PUSH(k, l, e1, ts) break;
      case TT_P_E: case TT_P_P:
        psihat[ti] = s1l; rss[ti] = 'L'; psihat[l] = tsr; rss[l] = 'R'; PUSH(k, l, e1, s1) break;
      case TT_O_OP: {
        int s2 = n2s[tsl * M + s1l];
        PUSH(k, l, e1, s1)
        PUSH(tsl, k, 7, s2)
        break;
      }
      case TT_E_P: {
        int s2 = n2s[tsl * M + s1l], s3 = n2s[s1r * M + tsr];
        int n1 = tj - l, n2 = k - ti;
        if (n1 == 0) { for (int p = ti; p < ti + n2; ++p) rss[p] = 'B'; }
        else if (n2 == 0) { for (int p = l; p < l + n1; ++p) rss[p] = 'B'; }
        else { for (int p = ti; p < ti + n2; ++p) rss[p] = 'I'; for (int p = l; p < l + n1; ++p) rss[p] = 'I'; }
        PUSH(l, tj, PL_L, s3)
        PUSH(ti, k, PL_L, s2)
        PUSH(k, l, e1, s1)
        break;
      }
      case TT_B_12: {
        int s2 = n2s[s1r * M + tsr];
        PUSH(l, tj, PL_2, s2)
        PUSH(k, l, e1, s1)
        break;
      }
      case TT_M_M: psihat[ti] = s1l; rss[ti] = 'M'; PUSH(k, l, PL_M, s1) break;
      default: break;
    }
  }
#undef PUSH
}

}  // namespace dp
}  // namespace relem
#endif
