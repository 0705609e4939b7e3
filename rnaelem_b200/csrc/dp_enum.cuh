// dp_enum.cuh -- the transition enumerators of the coupled grammar.
//
// One enumerator per parent state type.  Each visits, for one parent entry (cell (i,d), motif state s), every
// (structural transition x motif transition) term the reference would visit for that parent, IN THE REFERENCE'S
// VISITING ORDER (EnergyModel::compute_inside, energy_model.hpp:340-441, driving RNAelem::InsideFun,
// motif_model.hpp:230-423), and hands it to a visitor.  The same enumerators serve
//   * the inside pass        (visitor sums the terms in log space),
//   * the constrained passes (visitor vetoes terms: InsideEndFun / CYKFun, motif_scanner.hpp:581-665,802-913),
//   * the Viterbi pass       (visitor keeps the first strict maximum, motif_scanner.hpp:815-826),
//   * the outside pass       (visitor turns each term into a transition posterior and pushes it to the
//                             children: the top-down form of motif_trainer.hpp:330-458).
//
// Visitor concept:
//   bool  allow(const Emit&)                        veto by emission geometry (constraints)
//   void  t1(tt, c0, diff, tsc, slot, emit, geo)    one child in the band table at flat index c0
//   void  t2(tt, c0, c1, ...)                       two children (B <- 1 2)
//   void  t3(tt, c0, c1, c2, ...)                   three children (E <- P L L)
//   void  o1(tt, o0, ...)                           exterior: child O at exterior-table index o0
//   void  o2(tt, o0, c1, ...)                       exterior: children O (o0) and P (band index c1)
// diff = wt + lambda(s)*tsc exactly as motif_trainer.hpp:295 forms it.
#ifndef RELEM_DP_ENUM_CUH
#define RELEM_DP_ENUM_CUH
#include "dp_common.cuh"

namespace relem {
namespace dp {

struct ModelView {
  DevHMM h;
  DevParams p;
  DevEnergy en;
};

// emission geometry of a term: which sequence positions emit, and the parent / child motif states
struct Emit {
  int kind;    // 0 none, 1 pair (pos_l, pos_r), 2 right extension (pos_r), 3 left extension (pos_l)
  int pos_l, pos_r;
  int j;       // right end of the parent cell (needed by the "motif ends at L" rule)
  int sp, sc;  // parent / child state ids
};
struct Geo {
  int k, l, s1;  // what the reference stores in its Trace entry (motif_scanner.hpp:51-53)
};

RDEV double lam_of(const ModelView& m, int s, int& slot) {
  slot = (ld_ro(m.h.st_l + s) == ld_ro(m.h.st_r + s)) ? 0 : 1;
  return slot ? m.p.lambda1 : m.p.lambda0;
}
RDEV bool node_weighted(int c) { return c == '.' || c == '(' || c == ')'; }

// weight of a base pair emission (motif_model.hpp:271-300): parent s, child s1, bases at i and jm1
RDEV double pair_wt(const ModelView& m, const SeqView& q, int s, int s1, int i, int jm1) {
  const DevHMM& h = m.h;
  int sr = ld_ro(h.st_r + s), s1l = ld_ro(h.st_l + s1), s1r = ld_ro(h.st_r + s1);
  int xi = q.x[i], xj = q.x[jm1];
  int nr = ld_ro(h.node + sr), nl = ld_ro(h.node + s1l);
  double w = 0.;
  if (!m.p.no_prf) {
    if (nr == ')') {
      int t = bp_type(xi, xj);
      w = t ? ld_ro(m.p.theta + ld_ro(h.theta_off + ld_ro(h.theta_id + sr)) + t - 1) : 0.;
    } else {
      double a = xi ? ld_ro(m.p.theta + ld_ro(h.theta_off + ld_ro(h.theta_id + s1l)) + xi - 1) : 0.;
      double b = xj ? ld_ro(m.p.theta + ld_ro(h.theta_off + ld_ro(h.theta_id + sr)) + xj - 1) : 0.;
      w = a + b;
    }
  }
  double ws = (node_weighted(nl) ? q.ws[i] : 0.) + (node_weighted(nr) ? q.ws[jm1] : 0.);
  double t = (sr == s1r && ld_ro(h.node + s1r) == ')') ? m.p.ltau : 0.;
  return w + (t + ws);
}
// weight of a single-base emission by node hn at position pos; stay = self transition on a '.' node
RDEV double single_wt(const SeqView& q, int hn, int pos, bool stay) {
  return stay ? q.emitT[hn * q.L + pos] : q.emit0[hn * q.L + pos];
}

// ---- L(i,j,s) <- L(i,j-1,s1)   (InsideFun::before_transition, motif_model.hpp:243-257)
template <class V> RDEV void enum_L(const ModelView& m, const SeqView& q, int i, int d, int s, V& v) {
  if (d < 1 || !ld_ro(m.h.is_loop + s)) return;
  const DevHMM& h = m.h;
  int j = i + d, sr = ld_ro(h.st_r + s);
  bool dot = ld_ro(h.node + sr) == '.';
  for (int a = ld_ro(h.right_off + s), e = ld_ro(h.right_off + s + 1); a < e; ++a) {
    int s1 = ld_ro(h.right_idx + a);
    Emit em{2, -1, j - 1, j, s, s1};
    if (!v.allow(m, q, em)) continue;
    double wt = single_wt(q, sr, j - 1, dot && sr == ld_ro(h.st_r + s1));
    v.t1(TT_L_L, v.bidx(q, PL_L, i, d - 1, s1), wt, 0., 0, em, Geo{i, j - 1, s1});
  }
}

// ---- P(i,j,s) <- E(i+1,j-1,s1) | P(i+1,j-1,s1)
template <class V> RDEV void enum_P(const ModelView& m, const SeqView& q, int i, int d, int s, V& v) {
  const DevHMM& h = m.h;
  int j = i + d;
  int slot; double lam = lam_of(m, s, slot);
  int a0 = ld_ro(h.pair_off + s), a1 = ld_ro(h.pair_off + s + 1);
  if (a0 == a1) return;
  if (ok_E(q, i + 1, d - 2)) {
    for (int a = a0; a < a1; ++a) {
      int s1 = ld_ro(h.pair_idx + a);
      Emit em{1, i, j - 1, j, s, s1};
      if (!v.allow(m, q, em)) continue;
      double wt = pair_wt(m, q, s, s1, i, j - 1);
      v.t1(TT_P_E, v.bidx(q, PL_E, i + 1, d - 2, s1), wt, 0., slot, em, Geo{i + 1, j - 1, s1});
    }
  }
  if (ok_P(q, i + 1, d - 2)) {
    double tsc = m.en.no_ene ? 0. : e_loop(m.en, q, i, j - 1, i + 1, j - 2);
    if (tsc > NINF) {
      double lt = d_mul(lam, tsc);
      for (int a = a0; a < a1; ++a) {
        int s1 = ld_ro(h.pair_idx + a);
        Emit em{1, i, j - 1, j, s, s1};
        if (!v.allow(m, q, em)) continue;
        double wt = pair_wt(m, q, s, s1, i, j - 1);
        v.t1(TT_P_P, v.bidx(q, PL_P, i + 1, d - 2, s1), d_add(wt, lt), tsc, slot, em, Geo{i + 1, j - 1, s1});
      }
    }
  }
}

// ---- B(i,j,s) <- 1(i,k,(s.l,h)) 2(k,j,(h,s.r))
template <class V> RDEV void enum_B(const ModelView& m, const SeqView& q, int i, int d, int s, V& v) {
  const DevHMM& h = m.h;
  int j = i + d;
  int a0 = ld_ro(h.split_off + s), a1 = ld_ro(h.split_off + s + 1);
  Emit em{0, -1, -1, j, s, s};
  if (a0 == a1) return;
  if (q.lfr) {
    // same visiting order (k ascending), candidates found 32 at a time: bit u of row i of the left mask AND bit d-u of
    // row j of the right-indexed left mask
    const unsigned* ri = q.lf + i * q.mw;
    const unsigned* rj = q.lfr + j * q.mw;
    for (int u0 = 0; u0 <= d; u0 += 32) {
      int n = d - u0 + 1 < 32 ? d - u0 + 1 : 32;
      unsigned a = mask_window(ri, q.mw, u0, n);
      if (!a) continue;
      unsigned b = bit_rev(mask_window(rj, q.mw, d - u0 - 31, 32));   // bit t <-> position d - u0 - t
      unsigned mk = a & b;
      while (mk) {
        int t = bit_ffs(mk) - 1;
        mk &= mk - 1;
        int k = i + u0 + t;
        for (int aa = a0; aa < a1; ++aa) {
          int sl = ld_ro(h.split_left + aa), sr = ld_ro(h.split_right + aa);
          v.t2(TT_B_12, v.bidx(q, PL_1, i, k - i, sl), v.bidx(q, PL_2, k, j - k, sr), 0., 0., 0, em, Geo{i, k, sl});
        }
      }
    }
    return;
  }
  for (int k = i; k <= j; ++k) {
    if (!ok_B(q, i, k - i) || !ok_B(q, k, j - k)) continue;
    for (int a = a0; a < a1; ++a) {
      int sl = ld_ro(h.split_left + a), sr = ld_ro(h.split_right + a);
      v.t2(TT_B_12, v.bidx(q, PL_1, i, k - i, sl), v.bidx(q, PL_2, k, j - k, sr), 0., 0., 0, em, Geo{i, k, sl});
    }
  }
}

// ---- 2(i,j,s) <- 2(i,j-1,s1) | P(i,j,s)
template <class V> RDEV void enum_2(const ModelView& m, const SeqView& q, int i, int d, int s, V& v) {
  const DevHMM& h = m.h;
  int j = i + d;
  int slot; double lam = lam_of(m, s, slot);
  if (ok_B(q, i, d - 1)) {
    int sr = ld_ro(h.st_r + s);
    int nr = ld_ro(h.node + sr);
    for (int a = ld_ro(h.right_off + s), e = ld_ro(h.right_off + s + 1); a < e; ++a) {
      int s1 = ld_ro(h.right_idx + a);
      Emit em{2, -1, j - 1, j, s, s1};
      if (!v.allow(m, q, em)) continue;
      double wt = single_wt(q, sr, j - 1, nr == '.' && sr == ld_ro(h.st_r + s1));
      v.t1(TT_2_2, v.bidx(q, PL_2, i, d - 1, s1), wt, 0., slot, em, Geo{i, j - 1, s1});
    }
  }
  if (ok_P(q, i, d)) {
    double tsc = m.en.no_ene ? 0. : e_sum_ext_m(m.en, q, i, j - 1, false) + m.en.mlintern;
    if (tsc > NINF) {
      Emit em{0, -1, -1, j, s, s};
      v.t1(TT_2_P, v.bidx(q, PL_P, i, d, s), d_mul(lam, tsc), tsc, slot, em, Geo{i, j, s});
    }
  }
}

// ---- 1(i,j,s) <- 2(i,j,s) | B(i,j,s)
template <class V> RDEV void enum_1(const ModelView& m, const SeqView& q, int i, int d, int s, V& v) {
  int j = i + d;
  Emit em{0, -1, -1, j, s, s};
  v.t1(TT_1_2, v.bidx(q, PL_2, i, d, s), 0., 0., 0, em, Geo{i, j, s});
  v.t1(TT_1_B, v.bidx(q, PL_B, i, d, s), 0., 0., 0, em, Geo{i, j, s});
}

// ---- M(i,j,s) <- M(i+1,j,s1) | B(i,j,s)
template <class V> RDEV void enum_M(const ModelView& m, const SeqView& q, int i, int d, int s, V& v) {
  const DevHMM& h = m.h;
  int j = i + d;
  if (ok_M(q, i + 1, d - 1)) {
    int sl = ld_ro(h.st_l + s);
    bool dot = ld_ro(h.node + sl) == '.';
    for (int a = ld_ro(h.left_off + s), e = ld_ro(h.left_off + s + 1); a < e; ++a) {
      int s1 = ld_ro(h.left_idx + a);
      Emit em{3, i, -1, j, s, s1};
      if (!v.allow(m, q, em)) continue;
      int s1l = ld_ro(h.st_l + s1);
      double wt = single_wt(q, s1l, i, dot && sl == s1l);
      v.t1(TT_M_M, v.bidx(q, PL_M, i + 1, d - 1, s1), wt, 0., 0, em, Geo{i + 1, j, s1});
    }
  }
  if (ok_B(q, i, d)) {
    Emit em{0, -1, -1, j, s, s};
    v.t1(TT_M_B, v.bidx(q, PL_B, i, d, s), 0., 0., 0, em, Geo{i, j, s});
  }
}

// ---- E(i,j,s) <- M(i,j,s) | L(i,j,s) hairpin | P(k,l,s1) L(i,k,s2) L(l,j,s3)
template <class V> RDEV void enum_E(const ModelView& m, const SeqView& q, int i, int d, int s, V& v) {
  const DevHMM& h = m.h;
  int j = i + d;
  int slot; double lam = lam_of(m, s, slot);
  Emit em{0, -1, -1, j, s, s};
  if (ok_M(q, i, d)) {
    double tsc = m.en.no_ene ? 0. : e_sum_ext_m(m.en, q, j, i - 1, false) + (m.en.mlclosing + m.en.mlintern);
    if (tsc > NINF) v.t1(TT_E_M, v.bidx(q, PL_M, i, d, s), d_mul(lam, tsc), tsc, slot, em, Geo{i, j, s});
  }
  if (ld_ro(h.is_loop + s)) {
    double tsc = m.en.no_ene ? 0. : e_hairpin(m.en, q, i - 1, j);
    if (tsc > NINF) v.t1(TT_E_H, v.bidx(q, PL_L, i, d, s), d_mul(lam, tsc), tsc, slot, em, Geo{i, j, s});
  }
  int a0 = ld_ro(h.quad_off + s), a1 = ld_ro(h.quad_off + s + 1);
  if (a0 == a1) return;
  int C = q.C;
  // u1 <= C and u1 + u2 <= Cs.  Cs = C in the inside / Viterbi passes; the reference's outside pass does not bound u2
  // (energy_model.hpp:529), only the energy function does (30): that matters when --max-internal-loop is the binding
  // limit (see LinCtx::Csum in dp_lin.cuh)
  int Cs = C;
  if (V::kOutside && !m.en.no_ene && C < 30 && C < q.W - 7) Cs = 30;
  int lmin = j - Cs > i ? j - Cs : i;
  if (q.bpr) {
    // same visiting order (l descending, k ascending), inner pairs found by scanning row l of the right-indexed pair
    // mask: bit dd <-> pair (l-dd, l), k ascending = dd descending
    for (int l = j; l >= lmin; --l) {
      int kmax = i + Cs - (j - l);
      if (kmax > i + C) kmax = i + C;
      if (kmax > l) kmax = l;
      const unsigned* rl = q.bpr + l * q.mw;
      int dlo = l - kmax;
      for (int hi = l - i; hi >= dlo; hi -= 32) {
        int lo = hi - 31 > dlo ? hi - 31 : dlo;
        unsigned mk = mask_window(rl, q.mw, lo, hi - lo + 1);
        while (mk) {
          int t = bit_fls(mk) - 1;
          mk &= ~(1u << t);
          int k = l - (lo + t);
          if (k == i && l == j) continue;
          double tsc = m.en.no_ene ? 0. : e_loop(m.en, q, i - 1, j, k, l - 1);
          if (!(tsc > NINF)) continue;
          double lt = d_mul(lam, tsc);
          for (int a = a0; a < a1; ++a) {
            int s1 = ld_ro(h.quad_s1 + a), s2 = ld_ro(h.quad_s2 + a), s3 = ld_ro(h.quad_s3 + a);
            v.t3(TT_E_P, v.bidx(q, PL_P, k, l - k, s1), v.bidx(q, PL_L, i, k - i, s2),
                 v.bidx(q, PL_L, l, j - l, s3), lt, tsc, slot, em, Geo{k, l, s1});
          }
        }
      }
    }
    return;
  }
  for (int l = j; l >= lmin; --l) {
    int kmax = i + Cs - (j - l);
    if (kmax > i + C) kmax = i + C;
    if (kmax > l) kmax = l;
    for (int k = i; k <= kmax; ++k) {
      if (k == i && l == j) continue;
      if (!ok_P(q, k, l - k)) continue;
      double tsc = m.en.no_ene ? 0. : e_loop(m.en, q, i - 1, j, k, l - 1);
      if (!(tsc > NINF)) continue;
      double lt = d_mul(lam, tsc);
      for (int a = a0; a < a1; ++a) {
        int s1 = ld_ro(h.quad_s1 + a), s2 = ld_ro(h.quad_s2 + a), s3 = ld_ro(h.quad_s3 + a);
        v.t3(TT_E_P, v.bidx(q, PL_P, k, l - k, s1), v.bidx(q, PL_L, i, k - i, s2),
             v.bidx(q, PL_L, l, j - l, s3), lt, tsc, slot, em, Geo{k, l, s1});
      }
    }
  }
}

// ---- O(j,s) <- O(i,(s.l,h)) P(i,j,(h,s.r)) for i = j..i0, then O(j-1,s1)
template <class V> RDEV void enum_O(const ModelView& m, const SeqView& q, int j, int s, V& v) {
  const DevHMM& h = m.h;
  int slot; double lam = lam_of(m, s, slot);
  int a0 = ld_ro(h.split_off + s), a1 = ld_ro(h.split_off + s + 1);
  int i0 = j - q.W > 0 ? j - q.W : 0;
  Emit em0{0, -1, -1, j, s, s};
  if (a0 == a1) {
    // a state without splits takes no pair: nothing to visit
  } else if (q.bpr) {
    // same visiting order (i descending = span ascending), pairs closed at j found by scanning row j of the
    // right-indexed pair mask
    const unsigned* rj = q.bpr + j * q.mw;
    const int dmax = j - i0;
    for (int lo = 0; lo <= dmax; lo += 32) {
      unsigned mk = mask_window(rj, q.mw, lo, dmax - lo + 1 < 32 ? dmax - lo + 1 : 32);
      while (mk) {
        const int t = bit_ffs(mk) - 1;
        mk &= mk - 1;
        const int i = j - (lo + t);
        double tsc = m.en.no_ene ? 0. : e_sum_ext_m(m.en, q, i, j - 1, true);
        if (!(tsc > NINF)) continue;
        double lt = d_mul(lam, tsc);
        for (int a = a0; a < a1; ++a) {
          int sl = ld_ro(h.split_left + a), sr = ld_ro(h.split_right + a);
          v.o2(TT_O_OP, (unsigned)(i * q.S + sl), v.bidx(q, PL_P, i, j - i, sr), lt, tsc, slot, em0, Geo{i, j, sr});
        }
      }
    }
  } else
  for (int i = j; i >= i0; --i) {
    if (!ok_P(q, i, j - i)) continue;
    double tsc = m.en.no_ene ? 0. : e_sum_ext_m(m.en, q, i, j - 1, true);
    if (!(tsc > NINF)) continue;
    double lt = d_mul(lam, tsc);
    for (int a = a0; a < a1; ++a) {
      int sl = ld_ro(h.split_left + a), sr = ld_ro(h.split_right + a);
      v.o2(TT_O_OP, (unsigned)(i * q.S + sl), v.bidx(q, PL_P, i, j - i, sr), lt, tsc, slot, em0, Geo{i, j, sr});
    }
  }
  if (j > 0) {
    int sr = ld_ro(h.st_r + s);
    bool dot = ld_ro(h.node + sr) == '.';
    for (int a = ld_ro(h.right_off + s), e = ld_ro(h.right_off + s + 1); a < e; ++a) {
      int s1 = ld_ro(h.right_idx + a);
      Emit em{2, -1, j - 1, j, s, s1};
      if (!v.allow(m, q, em)) continue;
      double wt = single_wt(q, sr, j - 1, dot && sr == ld_ro(h.st_r + s1));
      v.o1(TT_O_O, (unsigned)((j - 1) * q.S + s1), wt, 0., slot, em, Geo{0, j - 1, s1});
    }
  }
}

// ------------------------------------------------------------------------------------------- constraints
// outside_cells(): how many leading / trailing cells of diagonal d can only hold one background state (see
// StartEndConstraint); 0 / 0 for constraints that cannot tell.
struct NoConstraint {
  RDEV bool ok(const ModelView&, const SeqView&, const Emit&) const { return true; }
  RDEV void outside_cells(int, int, int& nb, int& na) const { nb = 0; na = 0; }
};
// motif start fixed at Ys (InsideEndFun, motif_scanner.hpp:606-639; mirrored in OutsideEndFun :720-760)
struct StartConstraint {
  int ys;
  RDEV void outside_cells(int, int, int& nb, int& na) const { nb = 0; na = 0; }
  RDEV bool ok(const ModelView& m, const SeqView&, const Emit& e) const {
    if (e.kind == 0) return true;
    const DevHMM& h = m.h;
    if (e.kind == 1 || e.kind == 3) {
      if (e.pos_l == ys && !(ld_ro(h.st_l + e.sp) == 0 && ld_ro(h.st_l + e.sc) == 1)) return false;
    }
    if (e.kind == 1 || e.kind == 2) {
      if (e.pos_r == ys && !(ld_ro(h.st_r + e.sc) == 0 && ld_ro(h.st_r + e.sp) == 1)) return false;
    }
    return true;
  }
};
// motif start and end fixed (CYKFun, motif_scanner.hpp:843-880); ys = ye = -1 switches it off
struct StartEndConstraint {
  int ys, ye;
  // The node chain is monotone along the sequence and the vetoes below force the 0 -> 1 step onto position ys and the
  // M-2 -> M-1 step onto position ye.  A cell whose bases all lie before ys can therefore appear in a complete parse
  // only in state (0,0), one whose bases all lie after ye only in state (M-1,M-1); any other state of such a cell is
  // finite on its own but every parent that uses it runs into a veto.  The Viterbi pass skips those states (they are
  // stored as -inf): the root values, the arg-max path and its tie-breaking are unchanged, the work drops by ~S for
  // every cell outside the motif.  Margins of one position keep clear of the emission conventions at ys / ye.
  // cells of diagonal d are i = 0 .. ncell-1 with bases i .. i+d-1:  leading  i + d < ys,  trailing  i > ye + 1.
  RDEV void outside_cells(int d, int ncell, int& nb, int& na) const {
    nb = 0; na = 0;
    if (ys < 0 || ye < ys || d == 0) return;
    nb = ys - d; nb = nb < 0 ? 0 : (nb > ncell ? ncell : nb);
    na = ncell - (ye + 2); na = na < 0 ? 0 : na;
    if (na > ncell - nb) na = ncell - nb;
  }
  RDEV bool ok(const ModelView& m, const SeqView& q, const Emit& e) const {
    if (e.kind == 0) return true;
    const DevHMM& h = m.h;
    int M = h.M;
    int spl = ld_ro(h.st_l + e.sp), spr = ld_ro(h.st_r + e.sp);
    int scl = ld_ro(h.st_l + e.sc), scr = ld_ro(h.st_r + e.sc);
    if (e.kind == 1 || e.kind == 3) {
      if (e.pos_l == ys && !(spl == 0 && scl == 1)) return false;
    }
    if (e.kind == 1 || e.kind == 2) {
      if (e.pos_r == ys && !(scr == 0 && spr == 1)) return false;
    }
    if (e.kind == 1 || e.kind == 3) {
      if (e.pos_l == ye && !(spl == M - 2 && scl == M - 1)) return false;
    }
    if (e.kind == 1 || e.kind == 2) {
      if (e.pos_r == ye && !(scr == M - 2 && spr == M - 1)) return false;
      if (e.j == ye && q.L == e.j && spr != M - 2) return false;
    }
    return true;
  }
};

}  // namespace dp
}  // namespace relem
#endif
