// dp_warp.cuh -- warp-per-cell passes of the coupled grammar (the production mapping for S > 1).
//
// One warp owns one band cell (i,d) at a time (cells of a diagonal are handed out through a shared-memory
// counter).  Its 32 lanes run over the ENTRIES of the flattened motif-transition lists (right / left / pair
// transitions, (s.l,h,s.r) splits, loop-loop quads), so all lanes execute the same instruction stream while the
// structural candidates -- multiloop split points k, interior-loop inner pairs (k,l) -- are found once per warp
// with ballots over the bit masks and iterated uniformly.  Terms that share a target state sit in adjacent lanes
// (lists are grouped by target) and are combined with a segmented log-sum-exp reduction over warp shuffles; the
// per-cell values of all seven state types are staged in the warp's shared-memory slice, so the same-cell chain
// P -> 2 -> 1, B -> 1/M, {M,L} -> E never goes through global memory.  The outside pass walks the same lists
// top-down (see dp_pass.cuh) and pushes transition posteriors to the children with fp64 RED operations; same-cell
// children are owned by a lane and updated in shared memory without atomics.
#ifndef RELEM_DP_WARP_CUH
#define RELEM_DP_WARP_CUH
#include "dp_pass.cuh"
#include "dp_prim.cuh"

namespace relem {
namespace dp {

RDEV void lse_merge(Lse& a, double m2, double s2) {
  if (!(m2 > NINF)) return;
  if (!(a.m > NINF)) { a.m = m2; a.s = s2; }
  else if (m2 <= a.m) a.s += s2 * exp(m2 - a.m);
  else { a.s = a.s * exp(a.m - m2) + s2; a.m = m2; }
}

// per-warp shared-memory slice
struct WarpSm {
  double* cur;    // [NPLANE][S]  inside values of the cell being processed
  double* accm;   // [S] log-sum-exp accumulators (running max / scaled sum) per target state
  double* accs;
  double* qcur;   // [NPLANE][S][2] posteriors of the cell (outside pass)
  unsigned* kbits;  // valid split points of the cell
  int* pk; int* pl; double* pt;  // batch of structural candidates (k, l, energy)
  int S;
};
#define RELEM_PAIR_BATCH 64
RHD int warp_sm_bytes(int S, int Wmax) {
  int n = (NPLANE * S + 2 * S + NPLANE * S * 2) * 8 + ((Wmax + 64) / 32) * 4 + RELEM_PAIR_BATCH * 16;
  return (n + 15) & ~15;
}
RDEV WarpSm warp_carve(unsigned char* base, int S, int Wmax) {
  WarpSm w;
  w.S = S;
  w.cur = (double*)base; w.accm = w.cur + NPLANE * S; w.accs = w.accm + S; w.qcur = w.accs + S;
  w.pt = w.qcur + NPLANE * S * 2;
  w.pk = (int*)(w.pt + RELEM_PAIR_BATCH); w.pl = w.pk + RELEM_PAIR_BATCH;
  w.kbits = (unsigned*)(w.pl + RELEM_PAIR_BATCH);
  return w;
}

RDEV void acc_clear(WarpSm& w) {
  for (int t = lane_id(); t < w.S; t += WARP_N) { w.accm[t] = NINF; w.accs[t] = 0.; }
  w_sync();
}
RDEV double acc_value(const WarpSm& w, int t) { return w.accm[t] > NINF ? w.accm[t] + log(w.accs[t]) : NINF; }
// All lanes call.  key = target state of the lane's entry (entries of one target are adjacent), -1 = no entry.
RDEV void seg_commit(WarpSm& w, Lse v, int key) {
  const int lane = lane_id();
  for (int off = 1; off < WARP_N; off <<= 1) {
    double m2 = w_shfl_down(v.m, off), s2 = w_shfl_down(v.s, off);
    int k2 = w_shfl_down(key, off);
    if (lane + off < WARP_N && key >= 0 && k2 == key) lse_merge(v, m2, s2);
  }
  int kp = w_shfl_up(key, 1);
  if (key >= 0 && (lane == 0 || kp != key)) {
    Lse a; a.m = w.accm[key]; a.s = w.accs[key];
    lse_merge(a, v.m, v.s);
    w.accm[key] = a.m; w.accs[key] = a.s;
  }
  w_sync();
}

// valid multiloop split points u (k = i+u) of cell (i,d): both halves parsable (energy_model.hpp:359-363)
RDEV void find_splits(const SeqView& q, int i, int d, WarpSm& w) {
  for (int u0 = 0; u0 <= d; u0 += WARP_N) {
    int u = u0 + lane_id();
    bool ok = u <= d && ok_B(q, i, u) && ok_B(q, i + u, d - u);
    unsigned b = w_ballot(ok);
#if WARP_N == 32
    if (lane_id() == 0) w.kbits[u0 >> 5] = b;
#else
    if (b) w.kbits[u0 >> 5] |= 1u << (u0 & 31); else w.kbits[u0 >> 5] &= ~(1u << (u0 & 31));
#endif
  }
  w_sync();
}
#if WARP_N != 32
RDEV void clear_kbits(WarpSm& w, int d) { for (int t = 0; t <= d / 32; ++t) w.kbits[t] = 0u; }
#else
RDEV void clear_kbits(WarpSm&, int) {}
#endif

// Interior-loop candidates of E(i,j): inner pairs (k,l) with u1+u2 <= C (energy_model.hpp:413-426).  The walker
// fills the warp's batch buffer (k, l, loop energy) and calls `flush(n)` whenever it is full and at the end.
template <class F> RDEV void walk_inner_pairs(const ModelView& m, const SeqView& q, int i, int d, WarpSm& w, bool outside, F flush) {
  const int j = i + d, C = q.C, lane = lane_id();
  int n = 0;
  int Cs = C;   // bound on u1 + u2: see enum_E (dp_enum.cuh)
  if (outside && !m.en.no_ene && C < 30 && C < q.W - 7) Cs = 30;
  int lmin = j - Cs > i ? j - Cs : i;
  for (int l = j; l >= lmin; --l) {
    int kmax = i + Cs - (j - l);
    if (kmax > i + C) kmax = i + C;
    if (kmax > l) kmax = l;
    for (int k0 = i; k0 <= kmax; k0 += WARP_N) {
      int k = k0 + lane;
      bool ok = k <= kmax && !(k == i && l == j) && ok_P(q, k, l - k);
      double tsc = 0.;
      if (ok && !m.en.no_ene) { tsc = e_loop(m.en, q, i - 1, j, k, l - 1); ok = tsc > NINF; }
      unsigned bits = w_ballot(ok);
      while (bits) {
        int b = w_ffs(bits) - 1;
        bits &= bits - 1;
        double t = w_shfl(tsc, b);
        if (lane == 0) { w.pk[n] = k0 + b; w.pl[n] = l; w.pt[n] = t; }
        ++n;
        if (n == RELEM_PAIR_BATCH) { w_sync(); flush(n); n = 0; w_sync(); }
      }
    }
  }
  w_sync();
  if (n) flush(n);
  w_sync();
}
// exterior candidates of O(j): pairs (i,j) closed at j (energy_model.hpp:428-433)
template <class F> RDEV void walk_ext_pairs(const ModelView& m, const SeqView& q, int j, WarpSm& w, F flush) {
  const int lane = lane_id();
  int n = 0;
  int dmax = q.W < j ? q.W : j;
  for (int u0 = 0; u0 <= dmax; u0 += WARP_N) {
    int u = u0 + lane;
    int i = j - u;
    bool ok = u <= dmax && ok_P(q, i, u);
    double tsc = 0.;
    if (ok && !m.en.no_ene) { tsc = e_sum_ext_m(m.en, q, i, j - 1, true); ok = tsc > NINF; }
    unsigned bits = w_ballot(ok);
    while (bits) {
      int b = w_ffs(bits) - 1;
      bits &= bits - 1;
      double t = w_shfl(tsc, b);
      if (lane == 0) { w.pk[n] = j - (u0 + b); w.pl[n] = j; w.pt[n] = t; }
      ++n;
      if (n == RELEM_PAIR_BATCH) { w_sync(); flush(n); n = 0; w_sync(); }
    }
  }
  w_sync();
  if (n) flush(n);
  w_sync();
}

// ================================================================================================== inside
template <class CON>
RDEV void wc_inside_cell(const ModelView& m, const SeqView& q, double* tab, int i, int d, CON con, WarpSm& w) {
  const DevHMM& h = m.h;
  const int S = q.S, j = i + d, lane = lane_id();
  double* cur = w.cur;
  const bool gP = ok_P(q, i, d), gB = ok_B(q, i, d), gM = ok_M(q, i, d), gE = ok_E(q, i, d);
  // ---- L
  if (d == 0) {
    for (int s = lane; s < S; s += WARP_N) cur[PL_L * S + s] = (ld_ro(h.st_l + s) == ld_ro(h.st_r + s)) ? 0. : NINF;
  } else {
    acc_clear(w);
    for (int base = 0; base < h.n_right; base += WARP_N) {
      int a = base + lane, key = -1;
      Lse v; v.init();
      if (a < h.n_right) {
        int s = ld_ro(h.right_tgt + a);
        key = s;
        if (ld_ro(h.is_loop + s)) {
          int s1 = ld_ro(h.right_idx + a);
          Emit em{2, -1, j - 1, j, s, s1};
          if (con.ok(m, q, em)) {
            int sr = ld_ro(h.st_r + s);
            double wt = single_wt(q, sr, j - 1, ld_ro(h.node + sr) == '.' && sr == ld_ro(h.st_r + s1));
            v.add(tab[band_idx(q, PL_L, i, d - 1, s1)] + wt);
          }
        }
      }
      seg_commit(w, v, key);
    }
    for (int s = lane; s < S; s += WARP_N) cur[PL_L * S + s] = acc_value(w, s);
  }
  w_sync();
  // ---- P
  if (gP) {
    acc_clear(w);
    const bool cE = ok_E(q, i + 1, d - 2), cP = ok_P(q, i + 1, d - 2);
    double tsc = 0.;
    bool cPP = cP;
    if (cP && !m.en.no_ene) { tsc = e_loop(m.en, q, i, j - 1, i + 1, j - 2); cPP = tsc > NINF; }
    for (int base = 0; base < h.n_pair; base += WARP_N) {
      int a = base + lane, key = -1;
      Lse v; v.init();
      if (a < h.n_pair) {
        int s = ld_ro(h.pair_tgt + a), s1 = ld_ro(h.pair_idx + a);
        key = s;
        Emit em{1, i, j - 1, j, s, s1};
        if ((cE || cPP) && con.ok(m, q, em)) {
          double wt = pair_wt(m, q, s, s1, i, j - 1);
          if (cE) v.add(tab[band_idx(q, PL_E, i + 1, d - 2, s1)] + wt);
          if (cPP) { int slot; double lam = lam_of(m, s, slot); v.add(tab[band_idx(q, PL_P, i + 1, d - 2, s1)] + (wt + lam * tsc)); }
        }
      }
      seg_commit(w, v, key);
    }
    for (int s = lane; s < S; s += WARP_N) cur[PL_P * S + s] = acc_value(w, s);
    w_sync();
  }
  if (gB) {
    // ---- B
    clear_kbits(w, d);
    find_splits(q, i, d, w);
    acc_clear(w);
    const int nw = d / 32 + 1;
    for (int base = 0; base < h.n_split; base += WARP_N) {
      int a = base + lane, key = -1;
      Lse v; v.init();
      if (a < h.n_split) {
        key = ld_ro(h.split_tgt + a);
        int sl = ld_ro(h.split_left + a), sr = ld_ro(h.split_right + a);
        for (int t = 0; t < nw; ++t) {
          unsigned bits = w.kbits[t];
          while (bits) {
            int u = t * 32 + w_ffs(bits) - 1;
            bits &= bits - 1;
            v.add(tab[band_idx(q, PL_1, i, u, sl)] + tab[band_idx(q, PL_2, i + u, d - u, sr)]);
          }
        }
      }
      seg_commit(w, v, key);
    }
    for (int s = lane; s < S; s += WARP_N) cur[PL_B * S + s] = acc_value(w, s);
    w_sync();
    // ---- 2
    acc_clear(w);
    if (ok_B(q, i, d - 1)) {
      for (int base = 0; base < h.n_right; base += WARP_N) {
        int a = base + lane, key = -1;
        Lse v; v.init();
        if (a < h.n_right) {
          int s = ld_ro(h.right_tgt + a), s1 = ld_ro(h.right_idx + a);
          key = s;
          Emit em{2, -1, j - 1, j, s, s1};
          if (con.ok(m, q, em)) {
            int sr = ld_ro(h.st_r + s);
            double wt = single_wt(q, sr, j - 1, ld_ro(h.node + sr) == '.' && sr == ld_ro(h.st_r + s1));
            v.add(tab[band_idx(q, PL_2, i, d - 1, s1)] + wt);
          }
        }
        seg_commit(w, v, key);
      }
    }
    {
      double tsc = 0.;
      bool c2P = gP;
      if (gP && !m.en.no_ene) { tsc = e_sum_ext_m(m.en, q, i, j - 1, false) + m.en.mlintern; c2P = tsc > NINF; }
      for (int s = lane; s < S; s += WARP_N) {
        double x = acc_value(w, s);
        if (c2P) { int slot; double lam = lam_of(m, s, slot); x = lse2(x, cur[PL_P * S + s] + lam * tsc); }
        cur[PL_2 * S + s] = x;
        // ---- 1
        cur[PL_1 * S + s] = lse2(x, cur[PL_B * S + s]);
      }
    }
    w_sync();
  }
  // ---- M
  if (gM) {
    acc_clear(w);
    if (ok_M(q, i + 1, d - 1)) {
      for (int base = 0; base < h.n_left; base += WARP_N) {
        int a = base + lane, key = -1;
        Lse v; v.init();
        if (a < h.n_left) {
          int s = ld_ro(h.left_tgt + a), s1 = ld_ro(h.left_idx + a);
          key = s;
          Emit em{3, i, -1, j, s, s1};
          if (con.ok(m, q, em)) {
            int sl = ld_ro(h.st_l + s), s1l = ld_ro(h.st_l + s1);
            double wt = single_wt(q, s1l, i, ld_ro(h.node + sl) == '.' && sl == s1l);
            v.add(tab[band_idx(q, PL_M, i + 1, d - 1, s1)] + wt);
          }
        }
        seg_commit(w, v, key);
      }
    }
    for (int s = lane; s < S; s += WARP_N) {
      double x = acc_value(w, s);
      if (gB) x = lse2(x, cur[PL_B * S + s]);
      cur[PL_M * S + s] = x;
    }
    w_sync();
  }
  // ---- E
  if (gE) {
    acc_clear(w);
    if (h.n_quad > 0) {
      walk_inner_pairs(m, q, i, d, w, false, [&](int n) {
        for (int base = 0; base < h.n_quad; base += WARP_N) {
          int a = base + lane, key = -1;
          Lse v; v.init();
          if (a < h.n_quad) {
            int s = ld_ro(h.quad_tgt + a);
            key = s;
            int s1 = ld_ro(h.quad_s1 + a), s2 = ld_ro(h.quad_s2 + a), s3 = ld_ro(h.quad_s3 + a);
            int slot; double lam = lam_of(m, s, slot);
            for (int p = 0; p < n; ++p) {
              int k = w.pk[p], l = w.pl[p];
              double a0 = tab[band_idx(q, PL_P, k, l - k, s1)];
              if (!(a0 > NINF)) continue;
              v.add(a0 + (tab[band_idx(q, PL_L, i, k - i, s2)] + (tab[band_idx(q, PL_L, l, j - l, s3)] + lam * w.pt[p])));
            }
          }
          seg_commit(w, v, key);
        }
      });
    }
    double tM = 0., tH = 0.;
    bool cM = gM, cH = true;
    if (!m.en.no_ene) {
      if (gM) { tM = e_sum_ext_m(m.en, q, j, i - 1, false) + (m.en.mlclosing + m.en.mlintern); cM = tM > NINF; }
      tH = e_hairpin(m.en, q, i - 1, j); cH = tH > NINF;
    }
    for (int s = lane; s < S; s += WARP_N) {
      double x = acc_value(w, s);
      int slot; double lam = lam_of(m, s, slot);
      if (cM) x = lse2(x, cur[PL_M * S + s] + lam * tM);
      if (cH && ld_ro(h.is_loop + s)) x = lse2(x, cur[PL_L * S + s] + lam * tH);
      cur[PL_E * S + s] = x;
    }
    w_sync();
  }
  // ---- write back
  for (int s = lane; s < S; s += WARP_N) {
    tab[band_idx(q, PL_L, i, d, s)] = cur[PL_L * S + s];
    if (gP) tab[band_idx(q, PL_P, i, d, s)] = cur[PL_P * S + s];
    if (gB) {
      tab[band_idx(q, PL_B, i, d, s)] = cur[PL_B * S + s];
      tab[band_idx(q, PL_2, i, d, s)] = cur[PL_2 * S + s];
      tab[band_idx(q, PL_1, i, d, s)] = cur[PL_1 * S + s];
    }
    if (gM) tab[band_idx(q, PL_M, i, d, s)] = cur[PL_M * S + s];
    if (gE) tab[band_idx(q, PL_E, i, d, s)] = cur[PL_E * S + s];
  }
  w_sync();
}

// exterior recurrence, run by ONE warp (the caller keeps the other warps busy elsewhere)
template <class CON>
RDEV void wc_inside_ext(const ModelView& m, const SeqView& q, const double* tab, double* otab, CON con, WarpSm& w) {
  const DevHMM& h = m.h;
  const int S = q.S, L = q.L, lane = lane_id();
  for (int t = lane; t < (L + 1) * S; t += WARP_N) otab[t] = (t == h.s00) ? 0. : NINF;
  w_sync();
  for (int j = 1; j <= L; ++j) {
    acc_clear(w);
    walk_ext_pairs(m, q, j, w, [&](int n) {
      for (int base = 0; base < h.n_split; base += WARP_N) {
        int a = base + lane, key = -1;
        Lse v; v.init();
        if (a < h.n_split) {
          int s = ld_ro(h.split_tgt + a);
          key = s;
          int sl = ld_ro(h.split_left + a), sr = ld_ro(h.split_right + a);
          int slot; double lam = lam_of(m, s, slot);
          for (int p = 0; p < n; ++p) {
            int i = w.pk[p];
            v.add(otab[i * S + sl] + (tab[band_idx(q, PL_P, i, j - i, sr)] + lam * w.pt[p]));
          }
        }
        seg_commit(w, v, key);
      }
    });
    for (int base = 0; base < h.n_right; base += WARP_N) {
      int a = base + lane, key = -1;
      Lse v; v.init();
      if (a < h.n_right) {
        int s = ld_ro(h.right_tgt + a), s1 = ld_ro(h.right_idx + a);
        key = s;
        Emit em{2, -1, j - 1, j, s, s1};
        if (con.ok(m, q, em)) {
          int sr = ld_ro(h.st_r + s);
          double wt = single_wt(q, sr, j - 1, ld_ro(h.node + sr) == '.' && sr == ld_ro(h.st_r + s1));
          v.add(otab[(j - 1) * S + s1] + wt);
        }
      }
      seg_commit(w, v, key);
    }
    for (int s = lane; s < S; s += WARP_N) otab[j * S + s] = acc_value(w, s);
    w_sync();
  }
}

// ================================================================================================= outside
template <int NCH, int HOOK>
RDEV void emit_hooks(const ModelView& m, const SeqView& q, const Counts& cn, const Emit& e, const double* p) {
  if (HOOK == HOOK_NONE || e.kind == 0) return;
  const DevHMM& h = m.h;
  int spl = ld_ro(h.st_l + e.sp), spr = ld_ro(h.st_r + e.sp);
  int scl = ld_ro(h.st_l + e.sc), scr = ld_ro(h.st_r + e.sc);
  if ((HOOK == HOOK_TRAIN || HOOK == HOOK_SCAN_START) && !m.p.no_prf) {
    if (e.kind == 2) {
      for (int c = 0; c < NCH; ++c) red_add(cn.G + c * cn.ML + spr * q.L + e.pos_r, p[c]);
    } else if (e.kind == 3) {
      for (int c = 0; c < NCH; ++c) red_add(cn.G + c * cn.ML + scl * q.L + e.pos_l, p[c]);
    } else if (ld_ro(h.node + spr) == ')') {
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
  const int M = h.M;
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

// transition posterior for NCH channels: p[c] = q[c] * exp(val - in_y); returns false when nothing flows
template <int NCH> RDEV bool post(double val, double in_y, const double* qv, double* p) {
  if (!(val > NINF)) return false;
  double wgt = exp(val - in_y);
  bool any = false;
  for (int c = 0; c < NCH; ++c) { p[c] = qv[c] * wgt; any = any || p[c] != 0.; }
  return any;
}

template <int NCH, int HOOK, bool KEEP_P, class CON>
RDEV void wc_outside_cell(const ModelView& m, const SeqView& q, const double* tab, double* const* Q, int i, int d,
                          CON con, const Counts& cn, double* eh, WarpSm& w) {
  const DevHMM& h = m.h;
  const int S = q.S, j = i + d, lane = lane_id();
  double* cur = w.cur;
  double* qc = w.qcur;
#define QC(PL, s, c) qc[((PL) * S + (s)) * 2 + (c)]
  const bool gP = ok_P(q, i, d), gB = ok_B(q, i, d), gM = ok_M(q, i, d), gE = ok_E(q, i, d);
  bool gate[NPLANE];
  gate[PL_P] = gP; gate[PL_E] = gE; gate[PL_M] = gM; gate[PL_B] = gB; gate[PL_1] = gB; gate[PL_2] = gB; gate[PL_L] = true;
  for (int s = lane; s < S; s += WARP_N) {
    for (int pl = 0; pl < NPLANE; ++pl) {
      if (!gate[pl]) continue;
      unsigned idx = band_idx(q, pl, i, d, s);
      cur[pl * S + s] = tab[idx];
      for (int c = 0; c < NCH; ++c) QC(pl, s, c) = ld_cg(Q[c] + idx);
    }
  }
  w_sync();
  double p[NCH];
  // ---- E
  if (gE) {
    double tM = 0., tH = 0.;
    bool cM = gM, cH = true;
    if (!m.en.no_ene) {
      if (gM) { tM = e_sum_ext_m(m.en, q, j, i - 1, false) + (m.en.mlclosing + m.en.mlintern); cM = tM > NINF; }
      tH = e_hairpin(m.en, q, i - 1, j); cH = tH > NINF;
    }
    for (int s = lane; s < S; s += WARP_N) {
      double in_y = cur[PL_E * S + s];
      if (!(in_y > NINF)) continue;
      int slot; double lam = lam_of(m, s, slot);
      if (cM && post<NCH>(cur[PL_M * S + s] + lam * tM, in_y, &QC(PL_E, s, 0), p))
        for (int c = 0; c < NCH; ++c) { QC(PL_M, s, c) += p[c]; eh[c * 2 + slot] += tM * p[c]; }
      if (cH && ld_ro(h.is_loop + s) && post<NCH>(cur[PL_L * S + s] + lam * tH, in_y, &QC(PL_E, s, 0), p))
        for (int c = 0; c < NCH; ++c) { QC(PL_L, s, c) += p[c]; eh[c * 2 + slot] += tH * p[c]; }
    }
    if (h.n_quad > 0) {
      walk_inner_pairs(m, q, i, d, w, true, [&](int n) {
        for (int base = 0; base < h.n_quad; base += WARP_N) {
          int a = base + lane;
          if (a >= h.n_quad) continue;
          int s = ld_ro(h.quad_tgt + a);
          double in_y = cur[PL_E * S + s];
          if (!(in_y > NINF)) continue;
          bool anyq = false;
          for (int c = 0; c < NCH; ++c) anyq = anyq || QC(PL_E, s, c) != 0.;
          if (!anyq) continue;
          int s1 = ld_ro(h.quad_s1 + a), s2 = ld_ro(h.quad_s2 + a), s3 = ld_ro(h.quad_s3 + a);
          int slot; double lam = lam_of(m, s, slot);
          for (int pp = 0; pp < n; ++pp) {
            int k = w.pk[pp], l = w.pl[pp];
            unsigned c0 = band_idx(q, PL_P, k, l - k, s1);
            double a0 = tab[c0];
            if (!(a0 > NINF)) continue;
            unsigned c1 = band_idx(q, PL_L, i, k - i, s2), c2 = band_idx(q, PL_L, l, j - l, s3);
            double tsc = w.pt[pp];
            if (!post<NCH>(a0 + (tab[c1] + (tab[c2] + lam * tsc)), in_y, &QC(PL_E, s, 0), p)) continue;
            for (int c = 0; c < NCH; ++c) {
              red_add(Q[c] + c0, p[c]);
              if (k > i) red_add(Q[c] + c1, p[c]);
              if (l < j) red_add(Q[c] + c2, p[c]);
              eh[c * 2 + slot] += tsc * p[c];
            }
          }
        }
      });
    }
    w_sync();
  }
  // ---- M
  if (gM) {
    if (ok_M(q, i + 1, d - 1)) {
      for (int base = 0; base < h.n_left; base += WARP_N) {
        int a = base + lane;
        if (a >= h.n_left) continue;
        int s = ld_ro(h.left_tgt + a), s1 = ld_ro(h.left_idx + a);
        double in_y = cur[PL_M * S + s];
        if (!(in_y > NINF)) continue;
        Emit em{3, i, -1, j, s, s1};
        if (!con.ok(m, q, em)) continue;
        int sl = ld_ro(h.st_l + s), s1l = ld_ro(h.st_l + s1);
        double wt = single_wt(q, s1l, i, ld_ro(h.node + sl) == '.' && sl == s1l);
        unsigned c0 = band_idx(q, PL_M, i + 1, d - 1, s1);
        if (!post<NCH>(tab[c0] + wt, in_y, &QC(PL_M, s, 0), p)) continue;
        for (int c = 0; c < NCH; ++c) red_add(Q[c] + c0, p[c]);
        emit_hooks<NCH, HOOK>(m, q, cn, em, p);
      }
    }
    if (gB) {
      for (int s = lane; s < S; s += WARP_N) {
        double in_y = cur[PL_M * S + s];
        if (in_y > NINF && post<NCH>(cur[PL_B * S + s], in_y, &QC(PL_M, s, 0), p))
          for (int c = 0; c < NCH; ++c) QC(PL_B, s, c) += p[c];
      }
    }
    w_sync();
  }
  if (gB) {
    // ---- 1
    for (int s = lane; s < S; s += WARP_N) {
      double in_y = cur[PL_1 * S + s];
      if (!(in_y > NINF)) continue;
      if (post<NCH>(cur[PL_2 * S + s], in_y, &QC(PL_1, s, 0), p)) for (int c = 0; c < NCH; ++c) QC(PL_2, s, c) += p[c];
      if (post<NCH>(cur[PL_B * S + s], in_y, &QC(PL_1, s, 0), p)) for (int c = 0; c < NCH; ++c) QC(PL_B, s, c) += p[c];
    }
    w_sync();
    // ---- B
    clear_kbits(w, d);
    find_splits(q, i, d, w);
    const int nw = d / 32 + 1;
    for (int base = 0; base < h.n_split; base += WARP_N) {
      int a = base + lane;
      if (a >= h.n_split) continue;
      int s = ld_ro(h.split_tgt + a);
      double in_y = cur[PL_B * S + s];
      if (!(in_y > NINF)) continue;
      bool anyq = false;
      for (int c = 0; c < NCH; ++c) anyq = anyq || QC(PL_B, s, c) != 0.;
      if (!anyq) continue;
      int sl = ld_ro(h.split_left + a), sr = ld_ro(h.split_right + a);
      for (int t = 0; t < nw; ++t) {
        unsigned bits = w.kbits[t];
        while (bits) {
          int u = t * 32 + w_ffs(bits) - 1;
          bits &= bits - 1;
          unsigned c0 = band_idx(q, PL_1, i, u, sl), c1 = band_idx(q, PL_2, i + u, d - u, sr);
          if (!post<NCH>(tab[c0] + tab[c1], in_y, &QC(PL_B, s, 0), p)) continue;
          for (int c = 0; c < NCH; ++c) { red_add(Q[c] + c0, p[c]); red_add(Q[c] + c1, p[c]); }
        }
      }
    }
    // ---- 2
    if (ok_B(q, i, d - 1)) {
      for (int base = 0; base < h.n_right; base += WARP_N) {
        int a = base + lane;
        if (a >= h.n_right) continue;
        int s = ld_ro(h.right_tgt + a), s1 = ld_ro(h.right_idx + a);
        double in_y = cur[PL_2 * S + s];
        if (!(in_y > NINF)) continue;
        Emit em{2, -1, j - 1, j, s, s1};
        if (!con.ok(m, q, em)) continue;
        int sr = ld_ro(h.st_r + s);
        double wt = single_wt(q, sr, j - 1, ld_ro(h.node + sr) == '.' && sr == ld_ro(h.st_r + s1));
        unsigned c0 = band_idx(q, PL_2, i, d - 1, s1);
        if (!post<NCH>(tab[c0] + wt, in_y, &QC(PL_2, s, 0), p)) continue;
        for (int c = 0; c < NCH; ++c) red_add(Q[c] + c0, p[c]);
        emit_hooks<NCH, HOOK>(m, q, cn, em, p);
      }
    }
    if (gP) {
      double tsc = 0.;
      bool c2P = true;
      if (!m.en.no_ene) { tsc = e_sum_ext_m(m.en, q, i, j - 1, false) + m.en.mlintern; c2P = tsc > NINF; }
      if (c2P) {
        for (int s = lane; s < S; s += WARP_N) {
          double in_y = cur[PL_2 * S + s];
          if (!(in_y > NINF)) continue;
          int slot; double lam = lam_of(m, s, slot);
          if (post<NCH>(cur[PL_P * S + s] + lam * tsc, in_y, &QC(PL_2, s, 0), p))
            for (int c = 0; c < NCH; ++c) { QC(PL_P, s, c) += p[c]; eh[c * 2 + slot] += tsc * p[c]; }
        }
      }
    }
    w_sync();
  }
  // ---- P
  if (gP) {
    const bool cE = ok_E(q, i + 1, d - 2), cP = ok_P(q, i + 1, d - 2);
    double tsc = 0.;
    bool cPP = cP;
    if (cP && !m.en.no_ene) { tsc = e_loop(m.en, q, i, j - 1, i + 1, j - 2); cPP = tsc > NINF; }
    if (cE || cPP) {
      for (int base = 0; base < h.n_pair; base += WARP_N) {
        int a = base + lane;
        if (a >= h.n_pair) continue;
        int s = ld_ro(h.pair_tgt + a), s1 = ld_ro(h.pair_idx + a);
        double in_y = cur[PL_P * S + s];
        if (!(in_y > NINF)) continue;
        Emit em{1, i, j - 1, j, s, s1};
        if (!con.ok(m, q, em)) continue;
        double wt = pair_wt(m, q, s, s1, i, j - 1);
        if (cE) {
          unsigned c0 = band_idx(q, PL_E, i + 1, d - 2, s1);
          if (post<NCH>(tab[c0] + wt, in_y, &QC(PL_P, s, 0), p)) {
            for (int c = 0; c < NCH; ++c) red_add(Q[c] + c0, p[c]);
            emit_hooks<NCH, HOOK>(m, q, cn, em, p);
          }
        }
        if (cPP) {
          int slot; double lam = lam_of(m, s, slot);
          unsigned c0 = band_idx(q, PL_P, i + 1, d - 2, s1);
          if (post<NCH>(tab[c0] + (wt + lam * tsc), in_y, &QC(PL_P, s, 0), p)) {
            for (int c = 0; c < NCH; ++c) { red_add(Q[c] + c0, p[c]); eh[c * 2 + slot] += tsc * p[c]; }
            emit_hooks<NCH, HOOK>(m, q, cn, em, p);
          }
        }
      }
    }
    if (KEEP_P)
      for (int s = lane; s < S; s += WARP_N) Q[0][band_idx(q, PL_P, i, d, s)] = QC(PL_P, s, 0);
  }
  // ---- L
  if (d >= 1) {
    for (int base = 0; base < h.n_right; base += WARP_N) {
      int a = base + lane;
      if (a >= h.n_right) continue;
      int s = ld_ro(h.right_tgt + a);
      if (!ld_ro(h.is_loop + s)) continue;
      int s1 = ld_ro(h.right_idx + a);
      double in_y = cur[PL_L * S + s];
      if (!(in_y > NINF)) continue;
      Emit em{2, -1, j - 1, j, s, s1};
      if (!con.ok(m, q, em)) continue;
      int sr = ld_ro(h.st_r + s);
      double wt = single_wt(q, sr, j - 1, ld_ro(h.node + sr) == '.' && sr == ld_ro(h.st_r + s1));
      unsigned c0 = band_idx(q, PL_L, i, d - 1, s1);
      if (!post<NCH>(tab[c0] + wt, in_y, &QC(PL_L, s, 0), p)) continue;
      if (d > 1) for (int c = 0; c < NCH; ++c) red_add(Q[c] + c0, p[c]);
      emit_hooks<NCH, HOOK>(m, q, cn, em, p);
    }
  }
  w_sync();
#undef QC
}

// exterior, top-down, ONE warp.  QO must hold the root posteriors at j = L and zeros elsewhere.
template <int NCH, int HOOK, class CON>
RDEV void wc_outside_ext(const ModelView& m, const SeqView& q, const double* tab, const double* otab, double* const* Q,
                         double* const* QO, CON con, const Counts& cn, double* eh, WarpSm& w) {
  const DevHMM& h = m.h;
  const int S = q.S, L = q.L, lane = lane_id();
  double* cur = w.cur;
  double* qc = w.qcur;
  double p[NCH];
  for (int j = L; j >= 1; --j) {
    for (int s = lane; s < S; s += WARP_N) {
      cur[s] = otab[j * S + s];
      for (int c = 0; c < NCH; ++c) qc[s * 2 + c] = ld_cg(QO[c] + j * S + s);
    }
    w_sync();
    walk_ext_pairs(m, q, j, w, [&](int n) {
      for (int base = 0; base < h.n_split; base += WARP_N) {
        int a = base + lane;
        if (a >= h.n_split) continue;
        int s = ld_ro(h.split_tgt + a);
        double in_y = cur[s];
        if (!(in_y > NINF)) continue;
        bool anyq = false;
        for (int c = 0; c < NCH; ++c) anyq = anyq || qc[s * 2 + c] != 0.;
        if (!anyq) continue;
        int sl = ld_ro(h.split_left + a), sr = ld_ro(h.split_right + a);
        int slot; double lam = lam_of(m, s, slot);
        for (int pp = 0; pp < n; ++pp) {
          int i = w.pk[pp];
          double tsc = w.pt[pp];
          unsigned c1 = band_idx(q, PL_P, i, j - i, sr);
          if (!post<NCH>(otab[i * S + sl] + (tab[c1] + lam * tsc), in_y, qc + s * 2, p)) continue;
          for (int c = 0; c < NCH; ++c) {
            red_add(QO[c] + i * S + sl, p[c]);
            red_add(Q[c] + c1, p[c]);
            eh[c * 2 + slot] += tsc * p[c];
          }
        }
      }
    });
    for (int base = 0; base < h.n_right; base += WARP_N) {
      int a = base + lane;
      if (a >= h.n_right) continue;
      int s = ld_ro(h.right_tgt + a), s1 = ld_ro(h.right_idx + a);
      double in_y = cur[s];
      if (!(in_y > NINF)) continue;
      Emit em{2, -1, j - 1, j, s, s1};
      if (!con.ok(m, q, em)) continue;
      int sr = ld_ro(h.st_r + s);
      double wt = single_wt(q, sr, j - 1, ld_ro(h.node + sr) == '.' && sr == ld_ro(h.st_r + s1));
      if (!post<NCH>(otab[(j - 1) * S + s1] + wt, in_y, qc + s * 2, p)) continue;
      for (int c = 0; c < NCH; ++c) red_add(QO[c] + (j - 1) * S + s1, p[c]);
      emit_hooks<NCH, HOOK>(m, q, cn, em, p);
    }
    w_fence();
    w_sync();
  }
}

// =========================================================================================== CTA-level drivers
// ctr: two ints in shared memory (cell counters of alternating diagonals), both zero on entry and on exit.
template <class CON>
RDEV void cta_inside_warp(const ModelView& m, const SeqView& q, double* tab, double* otab, CON con, WarpSm& w, int* ctr,
                          double* zero_a, double* zero_b, unsigned long long zero_n) {
  const int L = q.L, W = q.W;
  for (int d = 0; d <= W; ++d) {
    int ncell = L + 1 - d;
    int* c = ctr + (d & 1);
    if (CTA_TID == 0) ctr[(d + 1) & 1] = 0;
    for (;;) {
      int cell = 0;
      if (lane_id() == 0) cell = ctr_next(c);
      cell = w_shfl(cell, 0);
      if (cell >= ncell) break;
      wc_inside_cell(m, q, tab, cell, d, con, w);
    }
    CTA_SYNC();
  }
  if (CTA_TID == 0) { ctr[0] = 0; ctr[1] = 0; }
  // exterior recurrence on warp 0; the other warps clear the posterior tables of the coming outside pass
  if (warp_id() == 0) wc_inside_ext(m, q, tab, otab, con, w);
#if WARP_N == 32
  if (n_warps() > 1) {
    if (warp_id() > 0)
      for (unsigned long long t = CTA_TID - 32; t < zero_n; t += CTA_NTH - 32) {
        zero_a[t] = 0.;
        if (zero_b) zero_b[t] = 0.;
      }
  } else
#endif
  {
    for (unsigned long long t = CTA_TID; t < zero_n; t += CTA_NTH) {
      zero_a[t] = 0.;
      if (zero_b) zero_b[t] = 0.;
    }
  }
  CTA_SYNC();
}

template <int NCH, int HOOK, bool KEEP_P, class CON>
RDEV void cta_outside_warp(const ModelView& m, const SeqView& q, const double* tab, const double* otab, double* Q0,
                           double* Q1, double* QO0, double* QO1, const double* root, CON con, Counts cn, double* eh_out,
                           WarpSm& w, int* ctr) {
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
  if (warp_id() == 0) wc_outside_ext<NCH, HOOK>(m, q, tab, otab, Qs, QOs, con, cn, eh, w);
  CTA_SYNC();
  for (int d = W; d >= 0; --d) {
    int ncell = L + 1 - d;
    int* c = ctr + (d & 1);
    if (CTA_TID == 0) ctr[(d + 1) & 1] = 0;
    for (;;) {
      int cell = 0;
      if (lane_id() == 0) cell = ctr_next(c);
      cell = w_shfl(cell, 0);
      if (cell >= ncell) break;
      wc_outside_cell<NCH, HOOK, KEEP_P>(m, q, tab, Qs, cell, d, con, cn, eh, w);
    }
    CTA_SYNC();
  }
  if (CTA_TID == 0) { ctr[0] = 0; ctr[1] = 0; }
  CTA_SYNC();
  for (int c = 0; c < NCH * 2; ++c) eh_out[c] = eh[c];
}

}  // namespace dp
}  // namespace relem
#endif
