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
  static constexpr bool kOutside = false;
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

// the same-cell chain L, P, B, 2, 1, M, E of one (cell, state) through the reference-order enumerators: used for the
// single-state cells outside the motif region, one thread per cell
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

// ---- cells that overlap the motif region (S states): one warp per cell, the 32 lanes over the ENTRIES of the flattened
// transition lists (right / left / pair transitions, splits, quads: DevHMM, grouped by parent state), structural
// candidates (split points, inner pairs) found once per warp from the bit masks.  Every candidate value is formed with
// the same explicit IEEE adds, in the same association, as MaxV forms it (dp_pass.cuh); max is exact, so the stored
// values are the reference's bit for bit although the order of evaluation is not.
#define RELEM_VIT_CAP 64
struct VitWarp {
  double* part;   // [n_max] per-entry maxima
  double* curB;   // [S] B of the cell
  double* bt;     // [CAP] loop energies of the buffered inner pairs
  int *bi, *bj;   // [CAP] buffered inner pairs (k, l)
  int* kbuf;      // [W+2] split points
};
RHD int vit_warp_bytes(int S, int Wmax, int n_max) {
  int n = (n_max + S + RELEM_VIT_CAP) * 8 + (2 * RELEM_VIT_CAP + Wmax + 4) * 4;
  return (n + 15) & ~15;
}
RDEV VitWarp vit_warp_carve(unsigned char* base, int S, int n_max) {
  VitWarp w;
  double* p = (double*)base;
  w.part = p; p += n_max;
  w.curB = p; p += S;
  w.bt = p; p += RELEM_VIT_CAP;
  int* ip = (int*)p;
  w.bi = ip; ip += RELEM_VIT_CAP;
  w.bj = ip; ip += RELEM_VIT_CAP;
  w.kbuf = ip;
  return w;
}
RDEV double vmax(double a, double b) { return a < b ? b : a; }
RDEV double vit_seg_max(const double* part, const int* off, int s) {
  const int a = ld_ro(off + s), n = ld_ro(off + s + 1) - a;
  double v = NINF;
  for (int k = 0; k < n; ++k) v = vmax(v, part[a + k]);
  return v;
}
// which single state a child cell keeps (-1: all of them)
RDEV int vit_only(const VitRegion& rg, int i, int d) {
  if (!rg.on || d == 0) return -1;
  if (i + d < rg.ys) return rg.s_bg0;
  if (i > rg.ye + 1) return rg.s_bgM;
  return -1;
}
RDEV double vit_ld(const SeqView& q, const double* tab, int plane, int i, int d, int s, int only) {
  return (only < 0 || s == only) ? tab[band_idx(q, plane, i, d, s)] : NINF;
}

template <class CON>
RDEV void vit_full_cell(const ModelView& m, const SeqView& q, double* tab, const CON& con, const VitRegion& rg, int i, int d,
                        VitWarp& w) {
  const DevHMM& h = m.h;
  const int S = q.S, j = i + d, lane = lane_id();
  const bool ne = m.en.no_ene != 0;
  if (d == 0) {
    for (int s = lane; s < S; s += WARP_N) tab[band_idx(q, PL_L, i, 0, s)] = (ld_ro(h.st_l + s) == ld_ro(h.st_r + s)) ? 0. : NINF;
    return;
  }
  // ---- L(i,j,s) <- L(i,j-1,s1) emitting x[j-1]     (enum_L)
  {
    const int on1 = vit_only(rg, i, d - 1);
    for (int a = lane; a < h.n_right; a += WARP_N) {
      const int s = ld_ro(h.right_tgt + a), s1 = ld_ro(h.right_idx + a);
      double v = NINF;
      if (ld_ro(h.is_loop + s)) {
        Emit em{2, -1, j - 1, j, s, s1};
        if (con.ok(m, q, em)) {
          const int sr = ld_ro(h.st_r + s);
          const double wt = single_wt(q, sr, j - 1, ld_ro(h.node + sr) == '.' && sr == ld_ro(h.st_r + s1));
          v = d_add(vit_ld(q, tab, PL_L, i, d - 1, s1, on1), wt);
        }
      }
      w.part[a] = v;
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N) tab[band_idx(q, PL_L, i, d, s)] = vit_seg_max(w.part, h.right_off, s);
    w_sync();
  }
  if (d < q.min_pair - 2) return;
  const bool gP = ok_P(q, i, d), gB = ok_B(q, i, d), gM = ok_M(q, i, d), gE = ok_E(q, i, d);
  // ---- P(i,j,s) <- E(i+1,j-1,s1) | P(i+1,j-1,s1)    (enum_P)
  if (gP) {
    const bool cE = ok_E(q, i + 1, d - 2), cP = ok_P(q, i + 1, d - 2);
    double tsc = 0.;
    bool cPP = cP;
    if (cP && !ne) { tsc = e_loop(m.en, q, i, j - 1, i + 1, j - 2); cPP = tsc > NINF; }
    const int on1 = vit_only(rg, i + 1, d - 2);
    for (int a = lane; a < h.n_pair; a += WARP_N) {
      const int s = ld_ro(h.pair_tgt + a), s1 = ld_ro(h.pair_idx + a);
      double v = NINF;
      Emit em{1, i, j - 1, j, s, s1};
      if ((cE || cPP) && con.ok(m, q, em)) {
        int slot; const double lam = lam_of(m, s, slot);
        const double wt = pair_wt(m, q, s, s1, i, j - 1);
        if (cE) v = vmax(v, d_add(vit_ld(q, tab, PL_E, i + 1, d - 2, s1, on1), wt));
        if (cPP) v = vmax(v, d_add(vit_ld(q, tab, PL_P, i + 1, d - 2, s1, on1), d_add(wt, d_mul(lam, tsc))));
      }
      w.part[a] = v;
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N) tab[band_idx(q, PL_P, i, d, s)] = vit_seg_max(w.part, h.pair_off, s);
    w_sync();
  }
  if (gB) {
    // ---- B(i,j,s) <- 1(i,k,sl) 2(k,j,sr)              (enum_B)
    int nk = 0;
    {
      const unsigned* ri = q.lf + i * q.mw;
      for (int u0 = 0; u0 <= d; u0 += WARP_N) {
        const int u = u0 + lane;
        const bool ok = u <= d && ((ri[u >> 5] >> (u & 31)) & 1u) && ok_B(q, i + u, d - u);
        const unsigned bal = w_ballot(ok);
        if (ok) w.kbuf[nk + w_popc(bal & lanemask_lt())] = u;
        nk += w_popc(bal);
      }
      w_sync();
    }
    for (int a = lane; a < h.n_split; a += WARP_N) {
      const int sl = ld_ro(h.split_left + a), sr = ld_ro(h.split_right + a);
      double v = NINF;
      for (int t = 0; t < nk; ++t) {
        const int u = w.kbuf[t];
        const int o1 = vit_only(rg, i, u), o2 = vit_only(rg, i + u, d - u);
        v = vmax(v, d_add(vit_ld(q, tab, PL_1, i, u, sl, o1), d_add(vit_ld(q, tab, PL_2, i + u, d - u, sr, o2), 0.)));
      }
      w.part[a] = v;
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N) {
      const double x = vit_seg_max(w.part, h.split_off, s);
      w.curB[s] = x;
      tab[band_idx(q, PL_B, i, d, s)] = x;
    }
    w_sync();
    // ---- 2(i,j,s) <- 2(i,j-1,s1) emitting x[j-1] | P(i,j,s);   1(i,j,s) <- 2(i,j,s) | B(i,j,s)     (enum_2, enum_1)
    const bool ok2 = ok_B(q, i, d - 1);
    const int on2 = vit_only(rg, i, d - 1);
    for (int a = lane; a < h.n_right; a += WARP_N) {
      const int s = ld_ro(h.right_tgt + a), s1 = ld_ro(h.right_idx + a);
      double v = NINF;
      Emit em{2, -1, j - 1, j, s, s1};
      if (ok2 && con.ok(m, q, em)) {
        const int sr = ld_ro(h.st_r + s);
        const double wt = single_wt(q, sr, j - 1, ld_ro(h.node + sr) == '.' && sr == ld_ro(h.st_r + s1));
        v = d_add(vit_ld(q, tab, PL_2, i, d - 1, s1, on2), wt);
      }
      w.part[a] = v;
    }
    double tsc2 = 0.;
    bool c2P = gP;
    if (gP && !ne) { tsc2 = e_sum_ext_m(m.en, q, i, j - 1, false) + m.en.mlintern; c2P = tsc2 > NINF; }
    w_sync();
    for (int s = lane; s < S; s += WARP_N) {
      double x = vit_seg_max(w.part, h.right_off, s);
      if (c2P) {
        int slot; const double lam = lam_of(m, s, slot);
        x = vmax(x, d_add(tab[band_idx(q, PL_P, i, d, s)], d_mul(lam, tsc2)));
      }
      tab[band_idx(q, PL_2, i, d, s)] = x;
      tab[band_idx(q, PL_1, i, d, s)] = vmax(d_add(x, 0.), d_add(w.curB[s], 0.));
    }
    w_sync();
  }
  // ---- M(i,j,s) <- M(i+1,j,s1) emitting x[i] | B(i,j,s)     (enum_M)
  if (gM) {
    const bool okM = ok_M(q, i + 1, d - 1);
    const int on1 = vit_only(rg, i + 1, d - 1);
    for (int a = lane; a < h.n_left; a += WARP_N) {
      const int s = ld_ro(h.left_tgt + a), s1 = ld_ro(h.left_idx + a);
      double v = NINF;
      Emit em{3, i, -1, j, s, s1};
      if (okM && con.ok(m, q, em)) {
        const int sl = ld_ro(h.st_l + s), s1l = ld_ro(h.st_l + s1);
        const double wt = single_wt(q, s1l, i, ld_ro(h.node + sl) == '.' && sl == s1l);
        v = d_add(vit_ld(q, tab, PL_M, i + 1, d - 1, s1, on1), wt);
      }
      w.part[a] = v;
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N) {
      double x = vit_seg_max(w.part, h.left_off, s);
      if (gB) x = vmax(x, d_add(w.curB[s], 0.));
      tab[band_idx(q, PL_M, i, d, s)] = x;
    }
    w_sync();
  }
  // ---- E(i,j,s) <- M(i,j,s) | L(i,j,s) hairpin | P(k,l,s1) L(i,k,s2) L(l,j,s3)     (enum_E)
  if (gE) {
    for (int a = lane; a < h.n_quad; a += WARP_N) w.part[a] = NINF;
    w_sync();
    if (h.n_quad > 0) {
      const int C = q.C, lo = d - C > 0 ? d - C : 0;
      int n = 0;
      auto flush = [&](int cnt) {
        // loop energies one inner pair per lane, then every quad entry over the buffered pairs
        w_sync();
        for (int z = lane; z < cnt; z += WARP_N) w.bt[z] = ne ? 0. : e_loop(m.en, q, i - 1, j, w.bi[z], w.bj[z] - 1);
        w_sync();
        for (int a = lane; a < h.n_quad; a += WARP_N) {
          const int s = ld_ro(h.quad_tgt + a), s1 = ld_ro(h.quad_s1 + a), s2 = ld_ro(h.quad_s2 + a), s3 = ld_ro(h.quad_s3 + a);
          int slot; const double lam = lam_of(m, s, slot);
          double v = w.part[a];
          for (int pp = 0; pp < cnt; ++pp) {
            const double tsc = w.bt[pp];
            if (!(tsc > NINF)) continue;
            const int k = w.bi[pp], l = w.bj[pp];
            const int oP = vit_only(rg, k, l - k), oL = vit_only(rg, i, k - i), oR = vit_only(rg, l, j - l);
            v = vmax(v, d_add(vit_ld(q, tab, PL_P, k, l - k, s1, oP),
                              d_add(vit_ld(q, tab, PL_L, i, k - i, s2, oL),
                                    d_add(vit_ld(q, tab, PL_L, l, j - l, s3, oR), d_mul(lam, tsc)))));
          }
          w.part[a] = v;
        }
        w_sync();
      };
      for (int u10 = 0; u10 <= C; u10 += WARP_N) {
        const int u1 = u10 + lane, k = i + u1;
        unsigned mk = 0u;
        if (u1 <= C && d - u1 >= lo) {
          mk = mask_window(q.bp + k * q.mw, q.mw, lo, d - u1 - lo + 1);
          if (u1 == 0) mk &= ~(1u << (d - lo));
        }
        while (w_any(mk != 0u)) {
          if (n > RELEM_VIT_CAP - WARP_N) { flush(n); n = 0; }
          const bool has = mk != 0u;
          int b = 0;
          if (has) { b = bit_ffs(mk) - 1; mk &= mk - 1; }
          const unsigned bal = w_ballot(has);
          if (has) { const int pos = n + w_popc(bal & lanemask_lt()); w.bi[pos] = k; w.bj[pos] = k + lo + b; }
          n += w_popc(bal);
        }
      }
      if (n) flush(n);
    }
    double tM = 0., tH = 0.;
    bool cM = gM, cH = true;
    if (!ne) {
      if (gM) { tM = e_sum_ext_m(m.en, q, j, i - 1, false) + (m.en.mlclosing + m.en.mlintern); cM = tM > NINF; }
      tH = e_hairpin(m.en, q, i - 1, j);
      cH = tH > NINF;
    }
    for (int s = lane; s < S; s += WARP_N) {
      int slot; const double lam = lam_of(m, s, slot);
      double x = vit_seg_max(w.part, h.quad_off, s);
      if (cM) x = vmax(x, d_add(tab[band_idx(q, PL_M, i, d, s)], d_mul(lam, tM)));
      if (cH && ld_ro(h.is_loop + s)) x = vmax(x, d_add(tab[band_idx(q, PL_L, i, d, s)], d_mul(lam, tH)));
      tab[band_idx(q, PL_E, i, d, s)] = x;
    }
    w_sync();
  }
}

// ---- a CTA works on a BATCH of sequences at once.  The band sweep needs one CTA barrier per diagonal, the exterior
// row is a serial recurrence over the positions and the traceback is one thread: with one sequence per CTA half of the
// warp time was spent waiting at those barriers (profiles/r2_viterbi.md).  With R sequences per CTA the units of a
// diagonal of all R sequences share one barrier, and exterior row and traceback of sequence r run on warp r alone
// (warp-level synchronisation only), R of them side by side.
struct VitSeq {
  SeqView q;
  double* tab;
  double* otab;
  int* stack;
  StartEndConstraint se;
  VitRegion rg;
  int n;          // sequence index in the batch view, -1 = empty place
  long long o;    // offset of its bases
};

template <class CON>
RDEV void cta_viterbi_band(const ModelView& m, const VitSeq* seqs, int nr, int Wmax, VitWarp& w) {
  const int lane = lane_id(), w0 = warp_id(), nw = n_warps();
  for (int d = 0; d <= Wmax; ++d) {
    int ubase = 0;   // units of this diagonal handed out so far (round robin over the warps, across the sequences)
    for (int r = 0; r < nr; ++r) {
      const VitSeq& z = seqs[r];
      if (z.n < 0 || d > z.q.W) continue;
      const SeqView q = z.q;   // private copy: the shared-memory original would be re-read at every use
      const int ncell = q.L + 1 - d;
      int nb, na;
      z.rg.cells(d, ncell, nb, na);
      const int nbg = nb + na, npack = (nbg + WARP_N - 1) / WARP_N, nfull = ncell - nbg;
      VitRegion off = z.rg;
      off.on = false;   // a single-state cell reads single-state entries only: no pruning test needed on its reads
      // static hand-out: the long packs first, the full cells follow.  Tried and dropped (profiles/r2_viterbi.md): units
      // through a shared atomic counter (wrong tables on the device, not understood); one warp per single-state cell
      // with lanes over candidates (correct, 1.5x slower than the packs).
      int u = (w0 - ubase % nw + nw) % nw;
      for (; u < npack + nfull; u += nw) {
        if (u < npack) {
          const int t = u * WARP_N + lane;
          if (t < nbg) {
            const bool lead = t < nb;
            vit_cell_state(m, q, z.tab, z.otab, (const CON&)z.se, off, lead ? t : ncell - na + (t - nb), d,
                           lead ? z.rg.s_bg0 : z.rg.s_bgM);
          }
        } else {
          vit_full_cell(m, q, z.tab, (const CON&)z.se, z.rg, nb + (u - npack), d, w);
        }
      }
      ubase += npack + nfull;
    }
    CTA_SYNC();
  }
}

// exterior recurrence of one sequence by ONE warp (never pruned: S values per position)
template <class CON> RDEV void warp_viterbi_exterior(const ModelView& m, const VitSeq& z) {
  const SeqView q = z.q;
  const int S = q.S, L = q.L, lane = lane_id();
  for (int t = lane; t < S; t += WARP_N) z.otab[t] = (t == m.h.s00) ? 0. : NINF;
  w_sync();
  for (int j = 1; j <= L; ++j) {
    for (int s = lane; s < S; s += WARP_N) {
      VitV<CON> v;
      v.tab = z.tab; v.otab = z.otab; v.con = (const CON&)z.se; v.rg = z.rg;
      v.init();
      enum_O(m, q, j, s, v);
      z.otab[j * S + s] = v.best;
    }
    w_sync();
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
