// dp_common.cuh -- device-side data model shared by the DP kernels (sm_100a).
//
// Layout in HBM (per sequence "slot", one CTA works on one slot at a time):
//   band tables   [plane e = P,E,M,B,1,2,L][i = 0..L][d = 0..W][s = 0..S-1]   fp64, s fastest
//                 (same index space as the reference's _inside[i][j-i][e][s], motif_trainer.hpp:62-71, but
//                 plane-major so that a row/column sweep of one state type touches only that plane)
//   exterior      [j = 0..L][s]                                               fp64
//   emit tables   [h = 0..M-1][p = 0..L-1] x {no tau, tau}                    fp64 (theta + position weight)
// Base-pair masks (bp_ok / left_bp_ok, energy_model.hpp:203-266) live in shared memory as bit rows.
//
// The file compiles in two modes: nvcc (the product) and, with RELEM_HOST_EMU defined, as plain C++ where one
// host thread plays a one-thread CTA.  The emulation exists only so that the kernel source can be debugged
// in a container without a GPU (tests/emu); librelem.so never contains it.
#ifndef RELEM_DP_COMMON_CUH
#define RELEM_DP_COMMON_CUH
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#ifdef RELEM_HOST_EMU
#define RDEV inline
#define RHD inline
#define RCONST static const
#define CTA_TID 0
#define CTA_NTH 1
#define CTA_SYNC() ((void)0)
template <class T> inline T ld_ro(const T* p) { return *p; }
template <class T> inline T ld_cg(const T* p) { return *p; }
inline void red_add(double* p, double v) { *p += v; }
inline double d_mul(double a, double b) { volatile double r = a * b; return r; }
inline double d_add(double a, double b) { volatile double r = a + b; return r; }
#else
#include <cuda_runtime.h>
#define RDEV __device__ __forceinline__
#define RHD __host__ __device__ inline
#define RCONST __constant__ const
#define CTA_TID ((int)threadIdx.x)
#define CTA_NTH ((int)blockDim.x)
#define CTA_SYNC() __syncthreads()
template <class T> RDEV T ld_ro(const T* p) { return __ldg(p); }
template <class T> RDEV T ld_cg(const T* p) { return __ldcg(p); }
RDEV void red_add(double* p, double v) { atomicAdd(p, v); }
// explicit IEEE ops: the Viterbi pass must reproduce the reference's sums bit for bit, so nothing on that
// path may be contracted into an FMA
RDEV double d_mul(double a, double b) { return __dmul_rn(a, b); }
RDEV double d_add(double a, double b) { return __dadd_rn(a, b); }
#endif

namespace relem {
namespace dp {

#ifdef RELEM_HOST_EMU
static const double NINF = -std::numeric_limits<double>::infinity();
#else
#define NINF (-CUDART_INF)
#endif
}  // namespace dp
}  // namespace relem

#ifndef RELEM_HOST_EMU
#include <math_constants.h>
#endif

namespace relem {
namespace dp {

// planes of the band table, in the reference's StateType order (energy_model.hpp:58-70)
enum { PL_P = 0, PL_E = 1, PL_M = 2, PL_B = 3, PL_1 = 4, PL_2 = 5, PL_L = 6, NPLANE = 7 };
// transition types, reference numbering (energy_model.hpp:72-91)
enum {
  TT_E_H = 0, TT_P_E, TT_P_P, TT_O_O, TT_O_OP, TT_E_P, TT_E_M, TT_M_M, TT_M_B, TT_B_12, TT_1_B, TT_1_2,
  TT_2_2, TT_2_P, TT_L_L, NTRANS
};

struct DevHMM {
  int M, S;
  const int *st_l, *st_r, *is_loop;
  const int *right_off, *right_idx, *left_off, *left_idx, *pair_off, *pair_idx;
  const int *quad_off, *quad_s1, *quad_s2, *quad_s3;
  const int *split_off, *split_left, *split_right;
  const int *node, *theta_id, *theta_off;
  const int *right_tgt, *left_tgt, *pair_tgt, *quad_tgt, *split_tgt;  // target state of each flat list entry
  int n_right, n_left, n_pair, n_quad, n_split;
  int s00, s0M2, s0M1;
};

struct DevEnergy {
  const double* hairpin_len;  // [max_span+2] length term incl. the >30 extrapolation
  const double* mismatch_h;   // [7][5][5]
  const double* mismatch_i;
  const double* mismatch_m;   // [8][5][5]
  const double* mismatch_1ni;
  const double* mismatch_23i;
  const double* mismatch_ext; // [8][5][5]
  const double* stack;        // [7][7]
  const double* bulge;        // [31]
  const double* internal;     // [31]
  const double* ninio;        // [31]
  const double* dangle5;      // [8][5]
  const double* dangle3;
  const double* int11;        // [8][8][5][5]
  const double* int21;        // [8][8][5][5][5]
  const double* int22;        // [8][8][5][5][5][5]
  double term_au, mlintern, mlclosing;
  // special hairpins: base-5 codes of the loop incl. closing pair, and their weights
  const int* tri_code; const double* tri_w; int ntri;
  const int* tetra_code; const double* tetra_w; int ntetra;
  const int* hexa_code; const double* hexa_w; int nhexa;
  int no_ene;
  int max_span, max_iloop;
  double min_lnbpp;  // log(min_bpp); -inf when min_bpp == 0
  int filter;        // min_bpp > 0
};

struct DevParams {
  const double* theta;  // flat rows
  int n_theta;
  double lambda0, lambda1, ltau;
  int no_prf;
};

// canonical pair types (bio_sequence.hpp:22-28): bp[a][b], 0 = not a pair
RCONST signed char BP_TYPE[25] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 5, 0, 0, 0, 1, 0, 0, 0, 2, 0, 3, 0, 6, 0, 4, 0};

RDEV int bp_type(int a, int b) { return BP_TYPE[a * 5 + b]; }

// ---- online log-sum-exp accumulator (replaces the pairwise log1p(exp()) of util.hpp:195-202: one exp per
// term, one log per table entry)
struct Lse {
  double m, s;
  RDEV void init() { m = NINF; s = 0.; }
  RDEV void add(double t) {
    if (!(t > NINF)) return;
    if (t <= m) s += exp(t - m);
    else { s = s * exp(m - t) + 1.; m = t; }
  }
  RDEV double value() const { return (m > NINF) ? m + log(s) : NINF; }
};

// per-sequence view used by all passes
struct SeqView {
  int L, W, C, W1;        // W1 = W+1
  int S;
  unsigned cells;         // (L+1)*(W+1)
  const unsigned char* x; // base codes, shared memory
  const unsigned* bp;     // bit rows [L+1][mw]
  const unsigned* lf;
  int mw;                 // words per mask row
  const signed char *sp3, *sp4, *sp6;  // special hairpin hit at position p (index or -1)
  const double* emit0;    // [M][L]
  const double* emitT;
  const double* ws;       // [L] (global)
  int min_pair;           // smallest pair span: turn+2 = 5
  int min_multi;          // smallest multiloop span: 2*(2+turn) = 10
  // optional (nullptr when absent): the same masks indexed by the RIGHT end of the span, bit d of row j <-> (j-d, j).
  // With them the enumerators find interior-loop and split candidates by bit scans instead of testing every (k,l).
  const unsigned* bpr;
  const unsigned* lfr;
};

RDEV bool mask_bit(const unsigned* rows, int mw, int i, int d) {
  return (rows[i * mw + (d >> 5)] >> (d & 31)) & 1u;
}
// gates of energy_model.hpp:289-338
RDEV bool ok_P(const SeqView& q, int i, int d) {
  return i >= 0 && d >= 0 && d <= q.W && i <= q.L && mask_bit(q.bp, q.mw, i, d);
}
RDEV bool ok_E(const SeqView& q, int i, int d) {
  return i > 0 && d >= 0 && d + 2 <= q.W && i <= q.L && mask_bit(q.bp, q.mw, i - 1, d + 2);
}
RDEV bool ok_M(const SeqView& q, int i, int d) { return i > 0 && i + d < q.L && d <= q.W && d >= q.min_multi; }
RDEV bool ok_B(const SeqView& q, int i, int d) {
  return i >= 0 && d >= 0 && d <= q.W && i <= q.L && mask_bit(q.lf, q.mw, i, d);
}
// n (1..32) bits of a mask row starting at bit lo (lo may be negative: bits below 0 read as 0); bits beyond the row
// read as 0
RDEV unsigned mask_window(const unsigned* row, int mw, int lo, int n) {
  int skip = 0;
  if (lo < 0) { skip = -lo; lo = 0; n -= skip; }
  if (n <= 0) return 0u;
  int w = lo >> 5, sh = lo & 31;
  unsigned a = w < mw ? row[w] : 0u, b = (w + 1) < mw ? row[w + 1] : 0u;
  unsigned v = sh ? ((a >> sh) | (b << (32 - sh))) : a;
  if (n < 32) v &= (1u << n) - 1u;
  return skip >= 32 ? 0u : (v << skip);
}
#ifdef RELEM_HOST_EMU
RDEV int bit_fls(unsigned v) { return v ? 32 - __builtin_clz(v) : 0; }       // 1-based index of the highest set bit
RDEV unsigned bit_rev(unsigned v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
}
RDEV int bit_ffs(unsigned v) { return __builtin_ffs((int)v); }
#else
RDEV int bit_fls(unsigned v) { return 32 - __clz((int)v); }
RDEV unsigned bit_rev(unsigned v) { return __brev(v); }
RDEV int bit_ffs(unsigned v) { return __ffs((int)v); }
#endif
RDEV unsigned band_idx(const SeqView& q, int plane, int i, int d, int s) {
  return ((unsigned)plane * q.cells + (unsigned)(i * q.W1 + d)) * (unsigned)q.S + (unsigned)s;
}

// ---- energies (energy_param.hpp:686-795), log-Boltzmann weights; i<j are base positions
RDEV double e_sum_ext_m(const DevEnergy& en, const SeqView& q, int i, int j, bool ext) {
  int type = bp_type(q.x[i], q.x[j]);
  double z = 0.;
  bool has5 = i - 1 >= 0, has3 = j + 1 < q.L;
  if (has5 && has3) {
    const double* t = ext ? en.mismatch_ext : en.mismatch_m;
    z = z + ld_ro(t + (type * 5 + q.x[i - 1]) * 5 + q.x[j + 1]);
    if (type > 2) z = z + en.term_au;
  } else {
    if (has5) z = z + ld_ro(en.dangle5 + type * 5 + q.x[i - 1]);
    if (has3) z = z + ld_ro(en.dangle3 + type * 5 + q.x[j + 1]);
    if (type > 2) z = z + en.term_au;
  }
  return z;
}

RDEV double e_hairpin(const DevEnergy& en, const SeqView& q, int i, int j) {
  int d = j - i - 1;
  if (d < 1) return NINF;
  int type = bp_type(q.x[i], q.x[j]);
  double z = ld_ro(en.hairpin_len + d);
  if (d < 3) {
  } else if (d == 3) {
    int hit = q.sp3[i];
    if (hit >= 0) return ld_ro(en.tri_w + hit);
    if (type > 2) z = z + en.term_au;
  } else if (d == 4) {
    int hit = q.sp4[i];
    if (hit >= 0) return ld_ro(en.tetra_w + hit);
  } else if (d == 6) {
    int hit = q.sp6[i];
    if (hit >= 0) return ld_ro(en.hexa_w + hit);
  }
  if (d > 3) z = z + ld_ro(en.mismatch_h + (type * 5 + q.x[i + 1]) * 5 + q.x[j - 1]);
  return z;
}

// closing pair (i,j), inner pair (p,q); i<p<q<j
RDEV double e_loop(const DevEnergy& en, const SeqView& sq, int i, int j, int p, int q) {
  const unsigned char* x = sq.x;
  int type = bp_type(x[i], x[j]);
  int type2 = bp_type(x[q], x[p]);
  int u1 = p - i - 1, u2 = j - q - 1;
  int u = u1 > u2 ? u1 : u2;
  double z;
  if (u1 < 0 || u2 < 0 || 30 < u1 + u2) return NINF;
  if (u1 == 0 && u2 == 0) return ld_ro(en.stack + type * 7 + type2);
  if (u1 == 0 || u2 == 0) {
    z = ld_ro(en.bulge + u);
    if (u == 1) z = z + ld_ro(en.stack + type * 7 + type2);
    else {
      if (type > 2) z = z + en.term_au;
      if (type2 > 2) z = z + en.term_au;
    }
    return z;
  }
  if (u <= 2) {
    if (u1 + u2 == 2) return ld_ro(en.int11 + ((type * 8 + type2) * 5 + x[i + 1]) * 5 + x[j - 1]);
    if (u1 == 1 && u2 == 2)
      return ld_ro(en.int21 + (((type * 8 + type2) * 5 + x[i + 1]) * 5 + x[q + 1]) * 5 + x[j - 1]);
    if (u1 == 2 && u2 == 1)
      return ld_ro(en.int21 + (((type2 * 8 + type) * 5 + x[q + 1]) * 5 + x[i + 1]) * 5 + x[p - 1]);
    return ld_ro(en.int22 + ((((type * 8 + type2) * 5 + x[i + 1]) * 5 + x[p - 1]) * 5 + x[q + 1]) * 5 + x[j - 1]);
  }
  int du = u1 - u2; if (du < 0) du = -du;
  z = ld_ro(en.internal + u1 + u2) + ld_ro(en.ninio + du);
  const double* mm = (u1 == 1 || u2 == 1) ? en.mismatch_1ni : (u1 + u2 == 5) ? en.mismatch_23i : en.mismatch_i;
  double a = ld_ro(mm + (type * 5 + x[i + 1]) * 5 + x[j - 1]);
  double b = ld_ro(mm + (type2 * 5 + x[q + 1]) * 5 + x[p - 1]);
  return z + (a + b);
}

}  // namespace dp
}  // namespace relem
#endif
