// dp_pass.cuh -- visitors and CTA-level passes built on the enumerators of dp_enum.cuh.
//
// Work decomposition: one CTA owns one sequence.  The band is swept as a wavefront over the span d = j-i; on a
// diagonal every (cell i, motif state s) pair is an independent work item handled by one thread, which walks the
// same-cell chain of state types itself (inside: L,P,B,2,1,M,E; outside: E,M,1,B,2,P,L), so the only
// synchronisation is one barrier per diagonal.  The exterior states O(j,.) form a 1-D recurrence over j that runs
// after (inside) / before (outside) the band sweep.
//
// Outside pass.  The reference pushes log-space outside values from parents to children and turns every visited
// transition into a posterior z = diff + inside(children) + outside(parent) - Z (motif_trainer.hpp:356-377).  Here
// the same posterior is formed top-down in linear space: Q(y) = posterior of parent entry y (known once all its
// parents are done), p(T) = Q(y) * exp(term(T) - inside(y)), and p(T) is added to Q of every child of T.  That
// needs exactly the inside enumeration again, so one enumerator serves both directions; all expected counts
// (EN, EH) and position posteriors are sums of p(T).  Posteriors below ~1e-308 flush to zero (the reference keeps
// them in log space); they cannot influence any count at fp64 resolution.
#ifndef RELEM_DP_PASS_CUH
#define RELEM_DP_PASS_CUH
#include "dp_enum.cuh"

namespace relem {
namespace dp {

// pairwise log-sum-exp exactly as util.hpp:195-202 (used only for the handful of partition-function sums)
RDEV double lse2(double x, double y) {
  if (!(y > NINF)) return x;
  if (!(x > NINF)) return y;
  return x < y ? y + log1p(exp(x - y)) : x + log1p(exp(y - x));
}

// ------------------------------------------------------------------------------------------------ visitors
template <class CON> struct SumV {
  static constexpr bool kOutside = false;
  RDEV unsigned bidx(const SeqView& q, int plane, int i, int d, int s) const { return band_idx(q, plane, i, d, s); }
  const double* tab;
  const double* otab;
  CON con;
  Lse acc;
  RDEV bool allow(const ModelView& m, const SeqView& q, const Emit& e) const { return con.ok(m, q, e); }
  RDEV void t1(int, unsigned c0, double diff, double, int, const Emit&, const Geo&) { acc.add(tab[c0] + diff); }
  RDEV void t2(int, unsigned c0, unsigned c1, double diff, double, int, const Emit&, const Geo&) {
    acc.add(tab[c0] + (tab[c1] + diff));
  }
  RDEV void t3(int, unsigned c0, unsigned c1, unsigned c2, double diff, double, int, const Emit&, const Geo&) {
    double a = tab[c0];
    if (!(a > NINF)) return;
    acc.add(a + (tab[c1] + (tab[c2] + diff)));
  }
  RDEV void o1(int, unsigned o0, double diff, double, int, const Emit&, const Geo&) { acc.add(otab[o0] + diff); }
  RDEV void o2(int, unsigned o0, unsigned c1, double diff, double, int, const Emit&, const Geo&) {
    acc.add(otab[o0] + (tab[c1] + diff));
  }
};

// Viterbi: first strict maximum in visiting order (CYKFun::compare, motif_scanner.hpp:815-826); sums are
// right-associated like util.hpp:223-224 and use explicit IEEE adds.
RDEV unsigned long long pack_trace(int tt, const Geo& g) {
  return ((unsigned long long)(unsigned)tt << 56) | ((unsigned long long)(unsigned)g.s1 << 40) |
         ((unsigned long long)(unsigned)g.k << 20) | (unsigned long long)(unsigned)g.l;
}
#define RELEM_NO_TRACE 0xFFFFFFFFFFFFFFFFull
template <class CON> struct MaxV {
  static constexpr bool kOutside = false;
  RDEV unsigned bidx(const SeqView& q, int plane, int i, int d, int s) const { return band_idx(q, plane, i, d, s); }
  const double* tab;
  const double* otab;
  CON con;
  double best;
  unsigned long long tr;
  RDEV void init() { best = NINF; tr = RELEM_NO_TRACE; }
  RDEV bool allow(const ModelView& m, const SeqView& q, const Emit& e) const { return con.ok(m, q, e); }
  RDEV void cmp(double y, int tt, const Geo& g) {
    if (best < y) { best = y; tr = pack_trace(tt, g); }
  }
  RDEV void t1(int tt, unsigned c0, double diff, double, int, const Emit&, const Geo& g) {
    cmp(d_add(tab[c0], diff), tt, g);
  }
  RDEV void t2(int tt, unsigned c0, unsigned c1, double diff, double, int, const Emit&, const Geo& g) {
    cmp(d_add(tab[c0], d_add(tab[c1], diff)), tt, g);
  }
  RDEV void t3(int tt, unsigned c0, unsigned c1, unsigned c2, double diff, double, int, const Emit&, const Geo& g) {
    cmp(d_add(tab[c0], d_add(tab[c1], d_add(tab[c2], diff))), tt, g);
  }
  RDEV void o1(int tt, unsigned o0, double diff, double, int, const Emit&, const Geo& g) {
    cmp(d_add(otab[o0], diff), tt, g);
  }
  RDEV void o2(int tt, unsigned o0, unsigned c1, double diff, double, int, const Emit&, const Geo& g) {
    cmp(d_add(otab[o0], d_add(tab[c1], diff)), tt, g);
  }
};

// what the outside pass accumulates besides Q
struct Counts {
  double* G;     // [NCH][M*L] single-base emission posteriors by (node, position)   (global, atomics)
  double* ENp;   // [NCH][n_theta] base-pair emission counts                          (shared, atomics)
  double* Pys;   // [L]   start posteriors (linear)   scan start pass                 (shared)
  double* Pyi;   // [L]   inner posteriors
  double* Pye;   // [L+1] end posteriors              scan end pass
  int n_theta;
  int ML;        // M*L
};
enum { HOOK_NONE = 0, HOOK_TRAIN = 1, HOOK_SCAN_START = 2, HOOK_SCAN_END = 3 };

RDEV int plane_of_same_cell(int tt) {
  switch (tt) {
    case TT_E_M: return PL_M;
    case TT_E_H: return PL_L;
    case TT_M_B: return PL_B;
    case TT_1_B: return PL_B;
    case TT_1_2: return PL_2;
    case TT_2_P: return PL_P;
  }
  return -1;
}

template <int NCH, int HOOK, class CON> struct ScatV {
  static constexpr bool kOutside = true;   // enum_E: the reference's outside pass visits a larger loop set
  RDEV unsigned bidx(const SeqView& q, int plane, int i, int d, int s) const { return band_idx(q, plane, i, d, s); }
  const double* tab;
  const double* otab;
  double* Q[NCH];
  double* QO[NCH];
  CON con;
  Counts cn;
  double in_y;
  double qy[NCH];
  double* loc;  // [NPLANE][NCH] same-cell contributions kept by the owning thread
  double* eh;   // [NCH][2]
  RDEV bool allow(const ModelView& m, const SeqView& q, const Emit& e) const { return con.ok(m, q, e); }

  RDEV void hooks(const ModelView& m, const SeqView& q, const Emit& e, const double* p) {
    if (HOOK == HOOK_NONE || e.kind == 0) return;
    const DevHMM& h = m.h;
    int spl = ld_ro(h.st_l + e.sp), spr = ld_ro(h.st_r + e.sp);
    int scl = ld_ro(h.st_l + e.sc), scr = ld_ro(h.st_r + e.sc);
    if ((HOOK == HOOK_TRAIN || HOOK == HOOK_SCAN_START) && !m.p.no_prf) {
      if (e.kind == 2) {
        for (int c = 0; c < NCH; ++c) red_add(cn.G + c * cn.ML + spr * q.L + e.pos_r, p[c]);
      } else if (e.kind == 3) {
        for (int c = 0; c < NCH; ++c) red_add(cn.G + c * cn.ML + scl * q.L + e.pos_l, p[c]);
      } else {
        if (ld_ro(h.node + spr) == ')') {
          int t = bp_type(q.x[e.pos_l], q.x[e.pos_r]);
          if (t > 0) {
            int idx = ld_ro(h.theta_off + ld_ro(h.theta_id + spr)) + t - 1;
            for (int c = 0; c < NCH; ++c) red_add(cn.ENp + c * cn.n_theta + idx, p[c]);
          }
        } else {
          for (int c = 0; c < NCH; ++c) {
            red_add(cn.G + c * cn.ML + scl * q.L + e.pos_l, p[c]);
            red_add(cn.G + c * cn.ML + spr * q.L + e.pos_r, p[c]);
          }
        }
      }
    }
    int M = h.M;
    if (HOOK == HOOK_SCAN_START) {
      if (e.kind == 1 || e.kind == 3) {
        if (spl == 0 && scl == 1) red_add(cn.Pys + e.pos_l, p[0]);
        if (scl != 0 && scl != M - 1) red_add(cn.Pyi + e.pos_l, p[0]);
      }
      if (e.kind == 1 || e.kind == 2) {
        if (scr == 0 && spr == 1) red_add(cn.Pys + e.pos_r, p[0]);
        if (spr != 0 && spr != M - 1) red_add(cn.Pyi + e.pos_r, p[0]);
      }
    }
    if (HOOK == HOOK_SCAN_END) {
      if (e.kind == 1 || e.kind == 3) {
        if (spl == M - 2 && scl == M - 1) red_add(cn.Pye + e.pos_l, p[0]);
      }
      if (e.kind == 1 || e.kind == 2) {
        if (scr == M - 2 && spr == M - 1) red_add(cn.Pye + e.pos_r, p[0]);
        if (spr == M - 2 && e.j == q.L) red_add(cn.Pye + q.L, p[0]);
      }
    }
  }
  RDEV bool weights(double val, double tsc, int slot, double* p) {
    if (!(val > NINF)) return false;
    double w = exp(val - in_y);
    bool any = false;
    for (int c = 0; c < NCH; ++c) {
      p[c] = qy[c] * w;
      any = any || (p[c] != 0.);
      eh[c * 2 + slot] += tsc * p[c];
    }
    return any;
  }
  RDEV void t1(int tt, unsigned c0, double diff, double tsc, int slot, const Emit& e, const Geo&) {
    // hooks need m,q: stored by the driver
    double p[NCH];
    if (!weights(tab[c0] + diff, tsc, slot, p)) return;
    int pl = plane_of_same_cell(tt);
    if (pl >= 0) { for (int c = 0; c < NCH; ++c) loc[pl * NCH + c] += p[c]; }
    else { for (int c = 0; c < NCH; ++c) red_add(Q[c] + c0, p[c]); }
    hooks(*mm, *qq, e, p);
  }
  RDEV void t2(int, unsigned c0, unsigned c1, double diff, double tsc, int slot, const Emit& e, const Geo&) {
    double p[NCH];
    if (!weights(tab[c0] + (tab[c1] + diff), tsc, slot, p)) return;
    for (int c = 0; c < NCH; ++c) { red_add(Q[c] + c0, p[c]); red_add(Q[c] + c1, p[c]); }
  }
  RDEV void t3(int, unsigned c0, unsigned c1, unsigned c2, double diff, double tsc, int slot, const Emit& e,
               const Geo&) {
    double a = tab[c0];
    if (!(a > NINF)) return;
    double p[NCH];
    if (!weights(a + (tab[c1] + (tab[c2] + diff)), tsc, slot, p)) return;
    for (int c = 0; c < NCH; ++c) { red_add(Q[c] + c0, p[c]); red_add(Q[c] + c1, p[c]); red_add(Q[c] + c2, p[c]); }
  }
  RDEV void o1(int, unsigned o0, double diff, double tsc, int slot, const Emit& e, const Geo&) {
    double p[NCH];
    if (!weights(otab[o0] + diff, tsc, slot, p)) return;
    for (int c = 0; c < NCH; ++c) red_add(QO[c] + o0, p[c]);
    hooks(*mm, *qq, e, p);
  }
  RDEV void o2(int, unsigned o0, unsigned c1, double diff, double tsc, int slot, const Emit& e, const Geo&) {
    double p[NCH];
    if (!weights(otab[o0] + (tab[c1] + diff), tsc, slot, p)) return;
    for (int c = 0; c < NCH; ++c) { red_add(QO[c] + o0, p[c]); red_add(Q[c] + c1, p[c]); }
  }
  const ModelView* mm;
  const SeqView* qq;
};

// ------------------------------------------------------------------------------------- per-sequence set-up
// canonical-pair mask (fill_bpp_tables, energy_model.hpp:213-219) into bit rows; returns nothing, counts later
RDEV void cta_clear_words(unsigned* w, int n) {
  for (int t = CTA_TID; t < n; t += CTA_NTH) w[t] = 0u;
}
RDEV void cta_canonical_mask(const SeqView& q, unsigned* bp) {
  // one thread per row word: no atomics needed
  int L = q.L, W = q.W, mw = q.mw;
  for (int t = CTA_TID; t < (L + 1) * mw; t += CTA_NTH) {
    int i = t / mw, w = t % mw;
    unsigned bits = 0u;
    for (int b = 0; b < 32; ++b) {
      int d = w * 32 + b;
      int j = i + d;
      if (d >= q.min_pair && d <= W && j <= L && bp_type(q.x[i], q.x[j - 1]) > 0) bits |= 1u << b;
    }
    bp[t] = bits;
  }
}
// left_bp_ok[i][d] = OR over d' <= d of bp_ok[i][d'] (fill_left_bpp_table, energy_model.hpp:203-209)
RDEV void cta_left_mask(const SeqView& q, const unsigned* bp, unsigned* lf) {
  int L = q.L, W = q.W, mw = q.mw;
  for (int i = CTA_TID; i <= L; i += CTA_NTH) {
    bool seen = false;
    for (int w = 0; w < mw; ++w) {
      unsigned in = bp[i * mw + w], out = 0u;
      if (seen) out = 0xFFFFFFFFu;
      else if (in) {
        int first = 0;
        while (!((in >> first) & 1u)) ++first;
        out = 0xFFFFFFFFu << first;
        seen = true;
      }
      // clip to d <= min(W, L-i)
      int dmax = W < L - i ? W : L - i;
      int lo = w * 32;
      if (dmax < lo) out = 0u;
      else if (dmax - lo < 31) out &= (0xFFFFFFFFu >> (31 - (dmax - lo)));
      lf[i * mw + w] = out;
    }
  }
}
RDEV int cta_count_bits(const unsigned* rows, int n, int* scratch) {
  // scratch: one int in shared memory
  if (CTA_TID == 0) *scratch = 0;
  CTA_SYNC();
  int c = 0;
  for (int t = CTA_TID; t < n; t += CTA_NTH) {
    unsigned v = rows[t];
    while (v) { v &= v - 1; ++c; }
  }
#ifdef RELEM_HOST_EMU
  *scratch += c;
#else
  if (c) atomicAdd(scratch, c);
#endif
  CTA_SYNC();
  int r = *scratch;
  CTA_SYNC();
  return r;
}

// special hairpin hits per start position (energy_param.hpp:723-737: the closing pair and the loop, matched
// against the list; first list entry wins)
RDEV void cta_special_hairpins(const DevEnergy& en, const unsigned char* x, int L, signed char* sp3,
                               signed char* sp4, signed char* sp6) {
  for (int p = CTA_TID; p < L; p += CTA_NTH) {
    int h3 = -1, h4 = -1, h6 = -1;
    if (p + 5 <= L && en.ntri) {
      int c = 0; for (int k = 0; k < 5; ++k) c = c * 5 + x[p + k];
      for (int a = 0; a < en.ntri; ++a) if (ld_ro(en.tri_code + a) == c) { h3 = a; break; }
    }
    if (p + 6 <= L && en.ntetra) {
      int c = 0; for (int k = 0; k < 6; ++k) c = c * 5 + x[p + k];
      for (int a = 0; a < en.ntetra; ++a) if (ld_ro(en.tetra_code + a) == c) { h4 = a; break; }
    }
    if (p + 8 <= L && en.nhexa) {
      int c = 0; for (int k = 0; k < 8; ++k) c = c * 5 + x[p + k];
      for (int a = 0; a < en.nhexa; ++a) if (ld_ro(en.hexa_code + a) == c) { h6 = a; break; }
    }
    sp3[p] = (signed char)h3; sp4[p] = (signed char)h4; sp6[p] = (signed char)h6;
  }
}

// emission tables: theta(h, x[p]) + position weight, without / with the self-loop penalty tau
// (motif_model.hpp:248-253: mulL(w, t, ws) = w + (t + ws))
RDEV void cta_emit_tables(const ModelView& m, const SeqView& q, double* emit0, double* emitT) {
  const DevHMM& h = m.h;
  int L = q.L;
  for (int t = CTA_TID; t < h.M * L; t += CTA_NTH) {
    int hn = t / L, p = t % L;
    int c = ld_ro(h.node + hn);
    double w = 0., ws = 0.;
    int tid = ld_ro(h.theta_id + hn);
    int b = q.x[p];
    if (!m.p.no_prf && tid >= 0 && b != 0 && (c == 'z' || c == '.' || c == '*' || c == 'o'))
      w = ld_ro(m.p.theta + ld_ro(h.theta_off + tid) + b - 1);
    if (node_weighted(c)) ws = q.ws[p];
    emit0[t] = w + (0. + ws);
    emitT[t] = w + (m.p.ltau + ws);
  }
}

// ----------------------------------------------------------------------------------------------- inside
// MAXMODE = false: log-sum-exp tables.  MAXMODE = true: Viterbi tables + packed trace entries.
template <bool MAXMODE, class CON>
RDEV void cta_inside(const ModelView& m, const SeqView& q, double* tab, double* otab, unsigned long long* trace,
                     unsigned long long* otrace, CON con) {
  const int S = q.S, L = q.L, W = q.W;
  const DevHMM& h = m.h;
  // background states of cells outside the motif (Viterbi with a fixed motif region only, see StartEndConstraint)
  int s_bg0 = h.s00, s_bgM = -1;
  if (MAXMODE)
    for (int s = 0; s < S; ++s)
      if (ld_ro(h.st_l + s) == h.M - 1 && ld_ro(h.st_r + s) == h.M - 1) s_bgM = s;
  for (int d = 0; d <= W; ++d) {
    int ncell = L + 1 - d;
    int nb = 0, na = 0;
    if (MAXMODE && s_bg0 >= 0 && s_bgM >= 0) con.outside_cells(d, ncell, nb, na);
    // states that cannot be part of a complete parse: -inf, no enumeration
    for (int t = CTA_TID; t < (nb + na) * S; t += CTA_NTH) {
      int c = t / S, s = t - c * S;
      int i = c < nb ? c : ncell - na + (c - nb);
      if (s == (c < nb ? s_bg0 : s_bgM)) continue;
      tab[band_idx(q, PL_L, i, d, s)] = NINF; trace[band_idx(q, PL_L, i, d, s)] = RELEM_NO_TRACE;
      if (ok_P(q, i, d)) { tab[band_idx(q, PL_P, i, d, s)] = NINF; trace[band_idx(q, PL_P, i, d, s)] = RELEM_NO_TRACE; }
      if (ok_B(q, i, d)) {
        tab[band_idx(q, PL_B, i, d, s)] = NINF; trace[band_idx(q, PL_B, i, d, s)] = RELEM_NO_TRACE;
        tab[band_idx(q, PL_2, i, d, s)] = NINF; trace[band_idx(q, PL_2, i, d, s)] = RELEM_NO_TRACE;
        tab[band_idx(q, PL_1, i, d, s)] = NINF; trace[band_idx(q, PL_1, i, d, s)] = RELEM_NO_TRACE;
      }
      if (ok_M(q, i, d)) { tab[band_idx(q, PL_M, i, d, s)] = NINF; trace[band_idx(q, PL_M, i, d, s)] = RELEM_NO_TRACE; }
      if (ok_E(q, i, d)) { tab[band_idx(q, PL_E, i, d, s)] = NINF; trace[band_idx(q, PL_E, i, d, s)] = RELEM_NO_TRACE; }
    }
    // everything else, compacted so that all lanes carry an enumeration: one state for each outside cell, S for the rest
    const int nwork = nb + na + (ncell - nb - na) * S;
    for (int t = CTA_TID; t < nwork; t += CTA_NTH) {
      int i, s;
      if (t < nb) { i = t; s = s_bg0; }
      else if (t < nb + na) { i = ncell - na + (t - nb); s = s_bgM; }
      else { int u = t - nb - na; i = nb + u / S; s = u - (u / S) * S; }
      // ---- L
      {
        unsigned idx = band_idx(q, PL_L, i, d, s);
        if (d == 0) {
          tab[idx] = (ld_ro(h.st_l + s) == ld_ro(h.st_r + s)) ? 0. : NINF;
          if (MAXMODE) trace[idx] = RELEM_NO_TRACE;
        } else if (MAXMODE) {
          MaxV<CON> v{tab, otab, con}; v.init();
          enum_L(m, q, i, d, s, v);
          tab[idx] = v.best; trace[idx] = v.tr;
        } else {
          SumV<CON> v{tab, otab, con}; v.acc.init();
          enum_L(m, q, i, d, s, v);
          tab[idx] = v.acc.value();
        }
      }
#define RELEM_RUN(PLANE, ENUM)                                                       \
  {                                                                                  \
    unsigned idx = band_idx(q, PLANE, i, d, s);                                      \
    if (MAXMODE) {                                                                   \
      MaxV<CON> v{tab, otab, con}; v.init();                                         \
      ENUM(m, q, i, d, s, v);                                                        \
      tab[idx] = v.best; trace[idx] = v.tr;                                          \
    } else {                                                                         \
      SumV<CON> v{tab, otab, con}; v.acc.init();                                     \
      ENUM(m, q, i, d, s, v);                                                        \
      tab[idx] = v.acc.value();                                                      \
    }                                                                                \
  }
      bool gP = ok_P(q, i, d), gB = ok_B(q, i, d), gM = ok_M(q, i, d), gE = ok_E(q, i, d);
      if (gP) RELEM_RUN(PL_P, enum_P)
      if (gB) {
        RELEM_RUN(PL_B, enum_B)
        RELEM_RUN(PL_2, enum_2)
        RELEM_RUN(PL_1, enum_1)
      }
      if (gM) RELEM_RUN(PL_M, enum_M)
      if (gE) RELEM_RUN(PL_E, enum_E)
#undef RELEM_RUN
        }
    CTA_SYNC();
  }
  // exterior recurrence
  for (int t = CTA_TID; t < (L + 1) * S; t += CTA_NTH) {
    otab[t] = (t == h.s00) ? 0. : NINF;
    if (MAXMODE) otrace[t] = RELEM_NO_TRACE;
  }
  CTA_SYNC();
  for (int j = 1; j <= L; ++j) {
    for (int s = CTA_TID; s < S; s += CTA_NTH) {
      if (MAXMODE) {
        MaxV<CON> v{tab, otab, con}; v.init();
        enum_O(m, q, j, s, v);
        otab[j * S + s] = v.best; otrace[j * S + s] = v.tr;
      } else {
        SumV<CON> v{tab, otab, con}; v.acc.init();
        enum_O(m, q, j, s, v);
        otab[j * S + s] = v.acc.value();
      }
    }
    CTA_SYNC();
  }
}

RDEV double part_func(const DevHMM& h, const double* otab, int L, int S, bool ari, bool nasi) {
  double a = (nasi && h.s00 >= 0) ? otab[L * S + h.s00] : NINF;
  double b = (ari && h.s0M2 >= 0) ? otab[L * S + h.s0M2] : NINF;
  double c = (ari && h.s0M1 >= 0) ? otab[L * S + h.s0M1] : NINF;
  return lse2(a, lse2(b, c));  // sumL(a, b, c) = sumL(a, sumL(b, c)), util.hpp:206
}

// ----------------------------------------------------------------------------------------------- outside
// Q tables must be zeroed by the caller.  root[c][3] = posterior of the three root exterior states
// {(0,0), (0,M-2), (0,M-1)} at j = L for channel c.  KEEP_P: write log Q of the P plane back (K0 needs it).
template <int NCH, int HOOK, class CON>
RDEV void cta_outside(const ModelView& m, const SeqView& q, const double* tab, const double* otab, double* Q0,
                      double* Q1, double* QO0, double* QO1, const double* root, CON con, Counts cn, double* eh_out) {
  const int S = q.S, L = q.L, W = q.W;
  const DevHMM& h = m.h;
  double eh[NCH * 2];
  for (int c = 0; c < NCH * 2; ++c) eh[c] = 0.;
  double* Qs[2] = {Q0, Q1};
  double* QOs[2] = {QO0, QO1};
  if (CTA_TID == 0) {
    for (int c = 0; c < NCH; ++c) {
      if (h.s00 >= 0) QOs[c][L * S + h.s00] = root[c * 3 + 0];
      if (h.s0M2 >= 0) QOs[c][L * S + h.s0M2] = root[c * 3 + 1];
      if (h.s0M1 >= 0) QOs[c][L * S + h.s0M1] = root[c * 3 + 2];
    }
  }
  CTA_SYNC();
  // exterior, top-down
  for (int j = L; j >= 1; --j) {
    for (int s = CTA_TID; s < S; s += CTA_NTH) {
      double in_y = otab[j * S + s];
      if (!(in_y > NINF)) continue;
      ScatV<NCH, HOOK, CON> v;
      v.tab = tab; v.otab = otab; v.con = con; v.cn = cn; v.in_y = in_y; v.loc = nullptr; v.eh = eh;
      v.mm = &m; v.qq = &q;
      bool any = false;
      for (int c = 0; c < NCH; ++c) {
        v.Q[c] = Qs[c]; v.QO[c] = QOs[c];
        v.qy[c] = ld_cg(QOs[c] + j * S + s);
        any = any || v.qy[c] != 0.;
      }
      if (any) enum_O(m, q, j, s, v);
    }
    CTA_SYNC();
  }
  // band, top-down
  for (int d = W; d >= 0; --d) {
    int ncell = L + 1 - d;
    for (int t = CTA_TID; t < ncell * S; t += CTA_NTH) {
      int i = t / S, s = t - i * S;
      double loc[NPLANE * NCH];
      for (int c = 0; c < NPLANE * NCH; ++c) loc[c] = 0.;
      ScatV<NCH, HOOK, CON> v;
      v.tab = tab; v.otab = otab; v.con = con; v.cn = cn; v.loc = loc; v.eh = eh;
      v.mm = &m; v.qq = &q;
      for (int c = 0; c < NCH; ++c) { v.Q[c] = Qs[c]; v.QO[c] = QOs[c]; }
#define RELEM_RUN(PLANE, ENUM)                                                 \
  {                                                                            \
    unsigned idx = band_idx(q, PLANE, i, d, s);                                \
    double in_y = tab[idx];                                                    \
    if (in_y > NINF) {                                                         \
      bool any = false;                                                        \
      for (int c = 0; c < NCH; ++c) {                                          \
        v.qy[c] = ld_cg(Qs[c] + idx) + loc[PLANE * NCH + c];                   \
        any = any || v.qy[c] != 0.;                                            \
      }                                                                        \
      if (PLANE == PL_P && loc[PL_P * NCH] != 0.) Qs[0][idx] = v.qy[0];        \
      if (any) { v.in_y = in_y; ENUM(m, q, i, d, s, v); }                      \
    }                                                                          \
  }
      bool gP = ok_P(q, i, d), gB = ok_B(q, i, d), gM = ok_M(q, i, d), gE = ok_E(q, i, d);
      if (gE) RELEM_RUN(PL_E, enum_E)
      if (gM) RELEM_RUN(PL_M, enum_M)
      if (gB) {
        RELEM_RUN(PL_1, enum_1)
        RELEM_RUN(PL_B, enum_B)
        RELEM_RUN(PL_2, enum_2)
      }
      if (gP) RELEM_RUN(PL_P, enum_P)
      if (d >= 1) RELEM_RUN(PL_L, enum_L)
#undef RELEM_RUN
    }
    CTA_SYNC();
  }
  for (int c = 0; c < NCH * 2; ++c) eh_out[c] = eh[c];
}

}  // namespace dp
}  // namespace relem
#endif
