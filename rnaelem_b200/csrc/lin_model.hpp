// lin_model.hpp -- host-side (C++14) preparation of what the scaled linear-space passes (dp_lin.cuh) read:
//   * the motif automaton's transition lists twice: grouped by PARENT state (inside pass: a target gathers from its
//     children) and grouped by CHILD state (outside pass: a child gathers from its parents), so that neither
//     direction needs atomics;
//   * emission factors exp(theta + tau) * kappa per list entry and base (RNAelem::InsideFun weights,
//     motif_model.hpp:243-300), and the theta slot every emission counts into (ProfileHMM::add_emit_count,
//     profile_hmm.hpp:144-186).
// Model set-up only; no DP runs on the host.
#ifndef RELEM_LIN_MODEL_HPP
#define RELEM_LIN_MODEL_HPP
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "host_model.hpp"

namespace relem {
namespace lin {

// entry flag bits beyond bit 0/1 (see r_flag / l_flag / p_flag below).  Single emissions use START/INNER/END(/ENDL);
// pair emissions use them for the left base and the R* set for the right base.
enum {
  LIN_F_START = 4, LIN_F_INNER = 8, LIN_F_END = 16, LIN_F_ENDL = 32,
  LIN_F_RSTART = 64, LIN_F_RINNER = 128, LIN_F_REND = 256, LIN_F_RENDL = 512
};

// device view of the automaton (all int32, one blob)
struct LinHMM {
  int M, S, n_right, n_left, n_pair, n_split, n_quad, n_max;
  int s00, s0M2, s0M1;
  const int *slot, *is_loop;
  const int *r_off, *r_tgt, *r_src, *r_flag;  // flag bit0: emitting node is position-weighted, bit1: target is a loop state
  const int *l_off, *l_tgt, *l_src, *l_flag;
  const int *p_off, *p_tgt, *p_src, *p_flag;  // bit0: left position weighted, bit1: right position weighted
  const int *sp_off, *sp_tgt, *sp_l, *sp_r;
  const int *q_off, *q_tgt, *q_s1, *q_s2, *q_s3;
  // child-grouped orders: X_ord[p] = entry index, X_coff = CSR over the child state
  const int *rT_off, *rT_ord, *lT_off, *lT_ord, *pT_off, *pT_ord;
  const int *spL_off, *spL_ord, *spR_off, *spR_ord;
  const int *qP_off, *qP_ord, *qL_off, *qL_ord, *qR_off, *qR_ord;
  const int *r_en, *l_en;      // [n][5]  theta slot of the emitted base, or -1
  const int *p_en1, *p_en2;    // [n][25] theta slots of a pair emission (x_left*5 + x_right), or -1
};

// per-parameter-set tables (doubles, one blob)
struct LinParams {
  const double *r_w, *l_w;  // [n][5]   linear emission factor incl. tau and one kappa
  const double *p_w;        // [n][25]  incl. tau and kappa^2
  double lambda0, lambda1;
  double kappa;             // per-base scale of the coupled tables
  double ln_kappa;
  int no_prf, n_theta;
};

struct LinHost {
  std::vector<int> blob;
  size_t o_slot, o_loop, o_roff, o_rtgt, o_rsrc, o_rflag, o_loff, o_ltgt, o_lsrc, o_lflag, o_poff, o_ptgt, o_psrc,
      o_pflag, o_spoff, o_sptgt, o_spl, o_spr, o_qoff, o_qtgt, o_q1, o_q2, o_q3, o_rToff, o_rTord, o_lToff, o_lTord,
      o_pToff, o_pTord, o_spLoff, o_spLord, o_spRoff, o_spRord, o_qPoff, o_qPord, o_qLoff, o_qLord, o_qRoff, o_qRord,
      o_ren, o_len, o_pen1, o_pen2;
  int M = 0, S = 0, n_right = 0, n_left = 0, n_pair = 0, n_split = 0, n_quad = 0, n_max = 0;
  int s00 = -1, s0M2 = -1, s0M1 = -1;

  size_t put(const std::vector<int>& v) {
    size_t o = blob.size();
    blob.insert(blob.end(), v.begin(), v.end());
    blob.push_back(0);
    return o;
  }
  // stable order of entries by child state; returns (CSR offsets over the child, order)
  static void by_child(const std::vector<int>& child, int S, std::vector<int>& off, std::vector<int>& ord) {
    ord.resize(child.size());
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return child[a] < child[b]; });
    off.assign(S + 1, 0);
    for (int c : child) off[c + 1] += 1;
    for (int s = 0; s < S; ++s) off[s + 1] += off[s];
  }
  static bool weighted(int c) { return c == '.' || c == '(' || c == ')'; }

  void build(const FlatHMM& f) {
    blob.clear();
    M = f.M; S = f.S;
    n_right = (int)f.right_idx.size(); n_left = (int)f.left_idx.size(); n_pair = (int)f.pair_idx.size();
    n_split = (int)f.split_left.size(); n_quad = (int)f.quad_s1.size();
    n_max = std::max(std::max(std::max(n_right, n_left), std::max(n_pair, n_split)), std::max(n_quad, S));
    s00 = f.s00; s0M2 = f.s0M2; s0M1 = f.s0M1;
    std::vector<int> slot(S);
    for (int s = 0; s < S; ++s) slot[s] = f.st_l[s] == f.st_r[s] ? 0 : 1;
    o_slot = put(slot); o_loop = put(f.is_loop);
    // ---- right: parent s emits x[j-1] with node s.r
    // scanner hooks (RNAelemScanDP::OutsideFun / OutsideEndFun, motif_scanner.hpp:420-800): which emissions mark the
    // motif start / an inner motif position / the motif end, by the parent's and child's interval ends
    std::vector<int> rflag(n_right), ren(n_right * 5, -1);
    for (int a = 0; a < n_right; ++a) {
      int s = f.right_tgt[a];
      int hn = f.st_r[s], c = f.node[hn];
      int spr = f.st_r[s], scr = f.st_r[f.right_idx[a]];
      rflag[a] = (weighted(c) ? 1 : 0) | (f.is_loop[s] ? 2 : 0) | ((scr == 0 && spr == 1) ? LIN_F_START : 0) |
                 ((spr != 0 && spr != M - 1) ? LIN_F_INNER : 0) | ((scr == M - 2 && spr == M - 1) ? LIN_F_END : 0) |
                 ((spr == M - 2) ? LIN_F_ENDL : 0);
      int tid = f.theta_id[hn];
      for (int b = 1; b < 5; ++b) if (tid >= 0) ren[a * 5 + b] = f.theta_off[tid] + b - 1;
    }
    o_roff = put(f.right_off); o_rtgt = put(f.right_tgt); o_rsrc = put(f.right_idx); o_rflag = put(rflag);
    // ---- left: parent s emits x[i] with node s1.l
    std::vector<int> lflag(n_left), len(n_left * 5, -1);
    for (int a = 0; a < n_left; ++a) {
      int s1 = f.left_idx[a];
      int hn = f.st_l[s1], c = f.node[hn];
      int spl = f.st_l[f.left_tgt[a]], scl = f.st_l[s1];
      lflag[a] = (weighted(c) ? 1 : 0) | ((spl == 0 && scl == 1) ? LIN_F_START : 0) |
                 ((scl != 0 && scl != M - 1) ? LIN_F_INNER : 0) | ((spl == M - 2 && scl == M - 1) ? LIN_F_END : 0);
      int tid = f.theta_id[hn];
      for (int b = 1; b < 5; ++b) if (tid >= 0) len[a * 5 + b] = f.theta_off[tid] + b - 1;
    }
    o_loff = put(f.left_off); o_ltgt = put(f.left_tgt); o_lsrc = put(f.left_idx); o_lflag = put(lflag);
    // ---- pair: parent s, child s1; emits x[i] with node s1.l and x[j-1] with node s.r
    static const int BP[25] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 5, 0, 0, 0, 1, 0, 0, 0, 2, 0, 3, 0, 6, 0, 4, 0};
    std::vector<int> pflag(n_pair), pen1(n_pair * 25, -1), pen2(n_pair * 25, -1);
    for (int a = 0; a < n_pair; ++a) {
      int s = f.pair_tgt[a], s1 = f.pair_idx[a];
      int sr = f.st_r[s], s1l = f.st_l[s1];
      int nr = f.node[sr], nl = f.node[s1l];
      int spl = f.st_l[s], scl = s1l, spr = sr, scr = f.st_r[s1];
      pflag[a] = (weighted(nl) ? 1 : 0) | (weighted(nr) ? 2 : 0) | ((spl == 0 && scl == 1) ? LIN_F_START : 0) |
                 ((scl != 0 && scl != M - 1) ? LIN_F_INNER : 0) | ((spl == M - 2 && scl == M - 1) ? LIN_F_END : 0) |
                 ((scr == 0 && spr == 1) ? LIN_F_RSTART : 0) | ((spr != 0 && spr != M - 1) ? LIN_F_RINNER : 0) |
                 ((scr == M - 2 && spr == M - 1) ? LIN_F_REND : 0) | ((spr == M - 2) ? LIN_F_RENDL : 0);
      for (int xi = 0; xi < 5; ++xi)
        for (int xj = 0; xj < 5; ++xj) {
          int k = a * 25 + xi * 5 + xj;
          if (nr == ')') {
            int t = BP[xi * 5 + xj], tid = f.theta_id[sr];
            if (t > 0 && tid >= 0) pen1[k] = f.theta_off[tid] + t - 1;
          } else {
            int t1 = f.theta_id[s1l], t2 = f.theta_id[sr];
            if (t1 >= 0 && xi != 0) pen1[k] = f.theta_off[t1] + xi - 1;
            if (t2 >= 0 && xj != 0) pen2[k] = f.theta_off[t2] + xj - 1;
          }
        }
    }
    o_poff = put(f.pair_off); o_ptgt = put(f.pair_tgt); o_psrc = put(f.pair_idx); o_pflag = put(pflag);
    o_spoff = put(f.split_off); o_sptgt = put(f.split_tgt); o_spl = put(f.split_left); o_spr = put(f.split_right);
    o_qoff = put(f.quad_off); o_qtgt = put(f.quad_tgt); o_q1 = put(f.quad_s1); o_q2 = put(f.quad_s2);
    o_q3 = put(f.quad_s3);
    std::vector<int> off, ord;
    by_child(f.right_idx, S, off, ord); o_rToff = put(off); o_rTord = put(ord);
    by_child(f.left_idx, S, off, ord); o_lToff = put(off); o_lTord = put(ord);
    by_child(f.pair_idx, S, off, ord); o_pToff = put(off); o_pTord = put(ord);
    by_child(f.split_left, S, off, ord); o_spLoff = put(off); o_spLord = put(ord);
    by_child(f.split_right, S, off, ord); o_spRoff = put(off); o_spRord = put(ord);
    by_child(f.quad_s1, S, off, ord); o_qPoff = put(off); o_qPord = put(ord);
    by_child(f.quad_s2, S, off, ord); o_qLoff = put(off); o_qLord = put(ord);
    by_child(f.quad_s3, S, off, ord); o_qRoff = put(off); o_qRord = put(ord);
    o_ren = put(ren); o_len = put(len); o_pen1 = put(pen1); o_pen2 = put(pen2);
  }

  void view(const int* b, LinHMM& d) const {
    d.M = M; d.S = S; d.n_right = n_right; d.n_left = n_left; d.n_pair = n_pair; d.n_split = n_split;
    d.n_quad = n_quad; d.n_max = n_max; d.s00 = s00; d.s0M2 = s0M2; d.s0M1 = s0M1;
    d.slot = b + o_slot; d.is_loop = b + o_loop;
    d.r_off = b + o_roff; d.r_tgt = b + o_rtgt; d.r_src = b + o_rsrc; d.r_flag = b + o_rflag;
    d.l_off = b + o_loff; d.l_tgt = b + o_ltgt; d.l_src = b + o_lsrc; d.l_flag = b + o_lflag;
    d.p_off = b + o_poff; d.p_tgt = b + o_ptgt; d.p_src = b + o_psrc; d.p_flag = b + o_pflag;
    d.sp_off = b + o_spoff; d.sp_tgt = b + o_sptgt; d.sp_l = b + o_spl; d.sp_r = b + o_spr;
    d.q_off = b + o_qoff; d.q_tgt = b + o_qtgt; d.q_s1 = b + o_q1; d.q_s2 = b + o_q2; d.q_s3 = b + o_q3;
    d.rT_off = b + o_rToff; d.rT_ord = b + o_rTord; d.lT_off = b + o_lToff; d.lT_ord = b + o_lTord;
    d.pT_off = b + o_pToff; d.pT_ord = b + o_pTord;
    d.spL_off = b + o_spLoff; d.spL_ord = b + o_spLord; d.spR_off = b + o_spRoff; d.spR_ord = b + o_spRord;
    d.qP_off = b + o_qPoff; d.qP_ord = b + o_qPord; d.qL_off = b + o_qLoff; d.qL_ord = b + o_qLord;
    d.qR_off = b + o_qRoff; d.qR_ord = b + o_qRord;
    d.r_en = b + o_ren; d.l_en = b + o_len; d.p_en1 = b + o_pen1; d.p_en2 = b + o_pen2;
  }
};

// Emission factor tables of one parameter set.  Layout of the returned blob: r_w [n_right*5], l_w [n_left*5],
// p_w [n_pair*25].  Weights follow cta_emit_tables / pair_wt of the log-space path (dp_pass.cuh, dp_enum.cuh), i.e.
// motif_model.hpp:243-300: theta only on 'z . * o' nodes for single emissions; the self-loop penalty tau on a '.'
// node that stays (right: s.r == s1.r, left: s.l == s1.l) and on a ')' node whose right end stays.
inline void build_lin_weights(const FlatHMM& f, const double* theta, double tau, int no_prf, double kappa,
                              std::vector<double>& blob, size_t& o_r, size_t& o_l, size_t& o_p) {
  static const int BP[25] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 5, 0, 0, 0, 1, 0, 0, 0, 2, 0, 3, 0, 6, 0, 4, 0};
  int n_right = (int)f.right_idx.size(), n_left = (int)f.left_idx.size(), n_pair = (int)f.pair_idx.size();
  blob.clear();
  auto single = [&](int hn, int b) {
    int c = f.node[hn], tid = f.theta_id[hn];
    double w = 0.;
    if (!no_prf && tid >= 0 && b != 0 && (c == 'z' || c == '.' || c == '*' || c == 'o')) w = theta[f.theta_off[tid] + b - 1];
    return w;
  };
  o_r = blob.size();
  for (int a = 0; a < n_right; ++a) {
    int s = f.right_tgt[a], s1 = f.right_idx[a];
    int sr = f.st_r[s];
    bool stay = f.node[sr] == '.' && sr == f.st_r[s1];
    for (int b = 0; b < 5; ++b) blob.push_back(std::exp(single(sr, b)) * (stay ? tau : 1.) * kappa);
  }
  o_l = blob.size();
  for (int a = 0; a < n_left; ++a) {
    int s = f.left_tgt[a], s1 = f.left_idx[a];
    int sl = f.st_l[s], s1l = f.st_l[s1];
    bool stay = f.node[sl] == '.' && sl == s1l;
    for (int b = 0; b < 5; ++b) blob.push_back(std::exp(single(s1l, b)) * (stay ? tau : 1.) * kappa);
  }
  o_p = blob.size();
  for (int a = 0; a < n_pair; ++a) {
    int s = f.pair_tgt[a], s1 = f.pair_idx[a];
    int sr = f.st_r[s], s1l = f.st_l[s1], s1r = f.st_r[s1];
    int nr = f.node[sr];
    bool stay = sr == s1r && f.node[s1r] == ')';
    for (int xi = 0; xi < 5; ++xi)
      for (int xj = 0; xj < 5; ++xj) {
        double w = 0.;
        if (!no_prf) {
          if (nr == ')') {
            int t = BP[xi * 5 + xj], tid = f.theta_id[sr];
            w = (t && tid >= 0) ? theta[f.theta_off[tid] + t - 1] : 0.;
          } else {
            int t1 = f.theta_id[s1l], t2 = f.theta_id[sr];
            double a1 = (xi && t1 >= 0) ? theta[f.theta_off[t1] + xi - 1] : 0.;
            double b1 = (xj && t2 >= 0) ? theta[f.theta_off[t2] + xj - 1] : 0.;
            w = a1 + b1;
          }
        }
        blob.push_back(std::exp(w) * (stay ? tau : 1.) * kappa * kappa);
      }
  }
}

}  // namespace lin
}  // namespace relem
#endif
