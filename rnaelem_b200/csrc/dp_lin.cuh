// dp_lin.cuh -- the throughput path of the E-step: scaled LINEAR-space inside / outside in gather form.
//
// Why it exists.  The log-space passes (dp_enum/dp_pass/dp_warp.cuh) follow the reference term by term: one
// exp per transition term, a log per table entry, posteriors pushed to children with fp64 atomics.  ncu showed
// that path bound by instruction count and latency (profiles/r1_estep_logspace.md).  The partition function is
// a polynomial in Boltzmann factors, so the same quantities can be computed with multiply-adds only:
//
//   * every table holds  a^(i,j,.) = exp(inside(i,j,.)) * kappa^(j-i)   with one per-base scale kappa chosen from
//     the background emission row (all recurrences are homogeneous in the span: children spans + emitted bases
//     add up to the parent span, so the scale never has to be tracked); the exterior row is scaled by kappa^j;
//   * outside values are kept as  b^(x) = d lnZ-weight / d a^(x)  (so posterior(x) = a^(x) * b^(x), scale free),
//     and every CHILD GATHERS from its parents (the reference scatters, motif_trainer.hpp:408-456): no atomics, results
//     are bit-reproducible from run to run.  The automaton's lists are therefore kept twice, grouped by parent and
//     grouped by child (lin_model.hpp);
//   * tables are stored per state type in the orientation their readers sweep: "by left end" [i][d][s] or "by right
//     end" [j][d][s] (multiloop splits read 1(i,.) along a row and 2(.,j) along a column; interior loops read the
//     inner pair by its right end and the two unpaired flanks one in each orientation);
//   * the difference ENo-ENx / EHo-EHx the trainer needs is linear in the root weights of the outside pass, so
//     one pass with root weights (1/Zo - [allowed]/Zx) yields it directly (NCH = 1); NCH = 2 keeps the two boundary
//     conditions apart for callers that ask for per-sequence detail.
//
// Work mapping: one CTA per sequence (persistent, queue), wavefront over the span d, one warp per band cell, the 32
// lanes over the entries of a transition list (or over structural candidates: split points, inner / outer pairs,
// found with bit-window extracts from the base-pair masks).  A sequence whose scaled values leave the fp64 range
// (non-finite Z or counts) is flagged and re-run by the caller on the log-space path.
//
// Reference semantics: EnergyModel::compute_inside/outside (energy_model.hpp:340-547), RNAelem::InsideFun /
// OutsideFun (motif_model.hpp:230-613), RNAelemTrainDP (motif_trainer.hpp:124-458), fill_bpp_tables
// (energy_model.hpp:211-266).
#ifndef RELEM_DP_LIN_CUH
#define RELEM_DP_LIN_CUH
#include "dp_prim.cuh"
#include "lin_model.hpp"

// unroll factor of the candidate loops of the interior-loop / exterior gathers (experiment switch; 1 = compiler default)
#if defined(LIN_PP_UNROLL) && LIN_PP_UNROLL == 2 && !defined(RELEM_HOST_EMU)
#define LIN_PP_PRAGMA _Pragma("unroll 2")
#elif defined(LIN_PP_UNROLL) && LIN_PP_UNROLL == 4 && !defined(RELEM_HOST_EMU)
#define LIN_PP_PRAGMA _Pragma("unroll 4")
#else
#define LIN_PP_PRAGMA
#endif
// split points a lane of the split gathers keeps in flight (2 or 4)
#ifndef LIN_SPLIT_UNROLL
#define LIN_SPLIT_UNROLL 4   // measured: 9 442 -> 9 532 sequence-evaluations/s at 64 registers, no spills
#endif

namespace relem {
namespace lin {
using namespace relem::dp;

#define LIN_CAP 64  // structural candidates buffered per warp before the list lanes consume them

// Interior loops in the outside pass.  1 (default): SCATTER -- once E(i,j) is final, the warp that owns it walks its
// inner pairs once and pushes the loop posteriors down to the inner pair P(k,l) and to the two unpaired flanks L(i,k),
// L(l,j) with fp64 RED operations (the reference's direction, motif_trainer.hpp:408-456); the loop energy is evaluated
// twice per E-step (inside E, outside E).  0: GATHER -- every child collects from its enclosing loops in three kernels
// of its own (outside P part 2, outside L parts 2 / 3): no atomics on tables, bit-reproducible, but the loop energy is
// evaluated four times and the enclosing-pair scans are repeated per child.  Measured: profiles/r2_scatter_ab.md.
#ifndef LIN_SCATTER_ILOOP
#define LIN_SCATTER_ILOOP 1
#endif

// Split gathers through shared memory with the bulk-copy engine (TMA, cp.async.bulk + mbarrier), A/B switch.
// 1: the [cell][state] vectors of up to LIN_TMA_STAGES split points are fetched by one 1-D bulk copy each into a
// per-warp staging buffer (no registers, no scoreboard: the whole batch is in flight at once), the lanes then take
// their operands from shared memory.  Needs 16-byte aligned vectors, i.e. an even number of states; other automata use
// the direct gathers.  0: direct per-lane gathers with four split points in flight.  Measured: profiles/r2_tma_ab.md.
#ifndef LIN_SPLIT_TMA
#define LIN_SPLIT_TMA 0
#endif
#ifndef LIN_TMA_STAGES
#define LIN_TMA_STAGES 16
#endif

struct LinEnergyScalars {  // linear-domain copies of the scalar energy terms
  double term_au, mlintern, mlclosing;
};

// model-wide data: __constant__ on the device (uploaded once per launch), so that the out-of-line energy
// functions below reach it without arguments
struct LinConst {
  LinHMM h;
  LinParams p;
  DevEnergy en;  // log-Boltzmann tables (coupled passes: exp(lambda * tsc), and tsc itself for the lambda gradient)
  DevEnergy el;  // the same tables exponentiated (energy-only filter pass, lambda = 1)
  double k0, k0sq;  // per-base scale of the energy-only pass
};
#ifdef RELEM_HOST_EMU
static LinConst LC;
#define LIN_NOINLINE __attribute__((noinline))
#else
__constant__ LinConst LC;
#define LIN_NOINLINE __device__ __noinline__
#endif

// per-sequence view, one copy per CTA in shared memory
struct LinCtx {
  SeqView q;     // x, special hairpins, bp (by left end), lf (by left end), sizes
  const unsigned* bpr;  // bp by right end: bit d of row j <-> pair (j-d, j)
  const unsigned* lfr;  // left_bp_ok by right end
  const double* wsf;    // exp(position weight) [L]
  const double* k0pow;  // kappa0^u, u = 0..W+1
  int Ceff;             // min(C, 30): loops longer than 30 have zero weight (energy_param.hpp:754-755)
  // The reference's OUTSIDE pass bounds the right flank of an interior loop by the loop variable itself
  // (energy_model.hpp:529), i.e. only by the energy function: it visits u1 <= C, u1 + u2 <= 30 where its inside pass
  // visits u1 + u2 <= C.  The sets differ only when --max-internal-loop is the binding limit (C < 30 and C < W - 7,
  // energies on); Csum / Cfl are then 30 / min(30, W), otherwise both equal Ceff (see lin_outside_limits).
  int Csum;             // bound on u1 + u2 in the outside pass
  int Cfl;              // longest flank the outside pass can reach
  // scanner: motif start fixed at position ys (-1: unconstrained), InsideEndFun / OutsideEndFun,
  // motif_scanner.hpp:606-639,720-760; linear start / inner / end posteriors [L+1] each
  int ys;
  double *pys, *pyi, *pye;
};

// scanner hooks (motif_scanner.hpp:420-579 start / inner, :667-800 end).  MODE 0: trainer, 1: start pass, 2: end pass
template <int MODE> RDEV void hook_emit(const LinCtx& c, int fl, int pos, double post) {
  if (MODE == 1) {
    if (fl & LIN_F_START) red_add(c.pys + pos, post);
    if (fl & LIN_F_INNER) red_add(c.pyi + pos, post);
  }
  if (MODE == 2) {
    if (fl & LIN_F_END) red_add(c.pye + pos, post);
  }
}
// right-hand emissions only: the motif may also end with the sequence
template <int MODE> RDEV void hook_emit_right(const LinCtx& c, int fl, int pos, double post) {
  hook_emit<MODE>(c, fl, pos, post);
  if (MODE == 2 && (fl & LIN_F_ENDL) && pos == c.q.L - 1) red_add(c.pye + c.q.L, post);
}

// flat offsets (one band table of one sequence has < 2^31 entries: checked by the host)
RDEV unsigned cidx(const SeqView& q, int row, int d) { return (unsigned)((row * q.W1 + d) * q.S); }
RDEV unsigned kidx(const SeqView& q, int row, int d) { return (unsigned)(row * q.W1 + d); }

// n (1..32) mask bits of a row starting at bit lo; bits beyond the row read as 0
RDEV unsigned win_bits(const unsigned* row, int mw, int lo, int n) {
  int w = lo >> 5, sh = lo & 31;
  unsigned a = w < mw ? row[w] : 0u, b = (w + 1) < mw ? row[w + 1] : 0u;
  unsigned v = sh ? ((a >> sh) | (b << (32 - sh))) : a;
  return n >= 32 ? v : (v & ((1u << n) - 1u));
}
RDEV bool row_bit(const unsigned* row, int d) { return (row[d >> 5] >> (d & 31)) & 1u; }

// ---- linear-domain energies (products of exponentiated tables; 0 = forbidden).  Same case analysis as
// e_sum_ext_m / e_hairpin / e_loop (dp_common.cuh), i.e. energy_param.hpp:686-795.
RDEV double l_sum_ext_m(const DevEnergy& el, const SeqView& q, int i, int j, bool ext) {
  int type = bp_type(q.x[i], q.x[j]);
  double z = 1.;
  bool has5 = i - 1 >= 0, has3 = j + 1 < q.L;
  if (has5 && has3) {
    const double* t = ext ? el.mismatch_ext : el.mismatch_m;
    z = ld_ro(t + (type * 5 + q.x[i - 1]) * 5 + q.x[j + 1]);
  } else {
    if (has5) z *= ld_ro(el.dangle5 + type * 5 + q.x[i - 1]);
    if (has3) z *= ld_ro(el.dangle3 + type * 5 + q.x[j + 1]);
  }
  if (type > 2) z *= el.term_au;
  return z;
}
RDEV double l_hairpin(const DevEnergy& el, const SeqView& q, int i, int j) {
  int d = j - i - 1;
  if (d < 1) return 0.;
  int type = bp_type(q.x[i], q.x[j]);
  double z = ld_ro(el.hairpin_len + d);
  if (d < 3) {
  } else if (d == 3) {
    int hit = q.sp3[i];
    if (hit >= 0) return ld_ro(el.tri_w + hit);
    if (type > 2) z *= el.term_au;
  } else if (d == 4) {
    int hit = q.sp4[i];
    if (hit >= 0) return ld_ro(el.tetra_w + hit);
  } else if (d == 6) {
    int hit = q.sp6[i];
    if (hit >= 0) return ld_ro(el.hexa_w + hit);
  }
  if (d > 3) z *= ld_ro(el.mismatch_h + (type * 5 + q.x[i + 1]) * 5 + q.x[j - 1]);
  return z;
}
RDEV double l_loop(const DevEnergy& el, const SeqView& sq, int i, int j, int p, int q) {
  const unsigned char* x = sq.x;
  int type = bp_type(x[i], x[j]);
  int type2 = bp_type(x[q], x[p]);
  int u1 = p - i - 1, u2 = j - q - 1;
  int u = u1 > u2 ? u1 : u2;
  if (u1 < 0 || u2 < 0 || 30 < u1 + u2) return 0.;
  if (u1 == 0 && u2 == 0) return ld_ro(el.stack + type * 7 + type2);
  if (u1 == 0 || u2 == 0) {
    double z = ld_ro(el.bulge + u);
    if (u == 1) z *= ld_ro(el.stack + type * 7 + type2);
    else {
      if (type > 2) z *= el.term_au;
      if (type2 > 2) z *= el.term_au;
    }
    return z;
  }
  if (u <= 2) {
    if (u1 + u2 == 2) return ld_ro(el.int11 + ((type * 8 + type2) * 5 + x[i + 1]) * 5 + x[j - 1]);
    if (u1 == 1 && u2 == 2)
      return ld_ro(el.int21 + (((type * 8 + type2) * 5 + x[i + 1]) * 5 + x[q + 1]) * 5 + x[j - 1]);
    if (u1 == 2 && u2 == 1)
      return ld_ro(el.int21 + (((type2 * 8 + type) * 5 + x[q + 1]) * 5 + x[i + 1]) * 5 + x[p - 1]);
    return ld_ro(el.int22 + ((((type * 8 + type2) * 5 + x[i + 1]) * 5 + x[p - 1]) * 5 + x[q + 1]) * 5 + x[j - 1]);
  }
  int du = u1 - u2; if (du < 0) du = -du;
  const double* mm = (u1 == 1 || u2 == 1) ? el.mismatch_1ni : (u1 + u2 == 5) ? el.mismatch_23i : el.mismatch_i;
  return ld_ro(el.internal + u1 + u2) * ld_ro(el.ninio + du) * ld_ro(mm + (type * 5 + x[i + 1]) * 5 + x[j - 1]) *
         ld_ro(mm + (type2 * 5 + x[q + 1]) * 5 + x[p - 1]);
}

// Out-of-line copies: the case analysis above is ~100 instructions and is needed in a dozen places; keeping one
// copy each keeps the kernels inside the instruction cache.
LIN_NOINLINE double nl_l_loop(const SeqView* q, int i, int j, int p, int qq) { return l_loop(LC.el, *q, i, j, p, qq); }
LIN_NOINLINE double nl_l_ext(const SeqView* q, int i, int j, int ext) { return l_sum_ext_m(LC.el, *q, i, j, ext != 0); }
LIN_NOINLINE double nl_l_hairpin(const SeqView* q, int i, int j) { return l_hairpin(LC.el, *q, i, j); }
LIN_NOINLINE double nl_e_loop(const SeqView* q, int i, int j, int p, int qq) { return e_loop(LC.en, *q, i, j, p, qq); }
LIN_NOINLINE double nl_e_ext(const SeqView* q, int i, int j, int ext) { return e_sum_ext_m(LC.en, *q, i, j, ext != 0); }
LIN_NOINLINE double nl_e_hairpin(const SeqView* q, int i, int j) { return e_hairpin(LC.en, *q, i, j); }
struct F2 { double f0, f1; };
// Boltzmann factors of a transition energy for the two lambda slots
LIN_NOINLINE F2 boltz2(double tsc) {
  F2 r;
  r.f0 = exp(LC.p.lambda0 * tsc);
  r.f1 = exp(LC.p.lambda1 * tsc);
  return r;
}

// Exterior rows run over the whole sequence, so a fixed per-base scale cannot keep them inside the fp64 range for long
// sequences.  They are stored as mantissas with one integer (power-of-two) exponent per position; a column is
// rescaled whenever its largest entry leaves [2^-256, 2^256].  Band tables need none of this: their span is <= W.
#ifndef LIN_RENORM_HI   // (tests/test_emu_parity.py also builds the emulation with tiny thresholds to exercise the rescaling)
#define LIN_RENORM_HI 1.157920892373162e77    // 2^256
#define LIN_RENORM_LO 8.636168555094445e-78   // 2^-256
#endif
RDEV int renorm_shift(double m) {  // m = largest magnitude of the column
  return (m > LIN_RENORM_HI || (m > 0. && m < LIN_RENORM_LO)) ? (int)ilogb(m) : 0;
}
#ifdef RELEM_HOST_EMU
RDEV double w_max(double v) { return v; }
#else
RDEV double w_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}
#endif

// ================================================================================= energy-only filter (K0)
// Tables of the energy-only grammar (one motif state, no emissions): EnergyModel::calc_BPP (energy_model.hpp:188-193).
// a: P (by right end), E, M, 1 (by left end), 2 (by right end); b: P, E, M, B (both orientations), 2.  L == 1.
struct K0Tabs {
  double *P, *E, *M, *o1, *o2, *O;
  double *bP, *bE, *bM, *bBl, *bBr, *b2, *bO;
  double* Pm;        // P(k,l) times the interior mismatch factor on the inner pair's side (by right end, like P)
  double* bEm;       // outside E(i,j) times the interior mismatch factor on the closing pair's side
  const double* G;   // [32][32] internal[u1+u2] * ninio[|u1-u2|] * kappa0^(u1+u2) for u1,u2 >= 3
  double *eO, *fO;   // [L+1] power-of-two exponents of the exterior rows O / bO (stored as doubles)
};

// Interior loops of the energy-only pass.  For u1,u2 >= 3 the loop energy is separable (energy_param.hpp:781-794:
// length term + asymmetry term + one mismatch term per closing pair, always from the generic mismatch table), so the
// sum over inner pairs is  mm(outer) * sum_{k,l} [P(k,l) mm(inner)] * G[u1][u2]  with the bracket stored once per
// pair: two loads and one FMA per candidate.  The remaining ("special": stack, bulges, 1x1, 1x2, 2x2, 1xn, 2x3) candidates
// are compacted across the warp and evaluated by the full case analysis, one per lane.
// sbuf: per-warp scratch of 128 ints.
RDEV void k0_specials_push(int* sbuf, int& ns, unsigned sp, int u1) {
  // exclusive prefix of the per-lane counts, then every lane writes its own candidates
  int cnt = w_popc(sp), pre = cnt;
#ifndef RELEM_HOST_EMU
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xFFFFFFFFu, pre, o);
    if ((int)(threadIdx.x & 31) >= o) pre += v;
  }
  int tot = __shfl_sync(0xFFFFFFFFu, pre, 31);
  pre -= cnt;
#else
  int tot = cnt;
  pre = 0;
#endif
  int pos = ns + pre;
  while (sp) {
    int b = w_ffs(sp) - 1;
    sp &= sp - 1;
    if (pos < 128) sbuf[pos] = (u1 << 8) | b;
    ++pos;
  }
  ns += tot;
}


RDEV void k0_inside_cell(const LinCtx& c, const K0Tabs& t, int i, int d, int* sbuf) {
  const SeqView& q = c.q;
  const DevEnergy& el = LC.el;
  const int j = i + d, lane = lane_id();
  const bool ne = el.no_ene != 0;
  const bool gP = ok_P(q, i, d), gB = ok_B(q, i, d), gM = ok_M(q, i, d), gE = ok_E(q, i, d);
  if (!(gP || gB || gM || gE)) return;
  double vP = 0.;
  if (gP) {
    if (ok_E(q, i + 1, d - 2)) vP += t.E[kidx(q, i + 1, d - 2)];
    if (ok_P(q, i + 1, d - 2)) vP += t.P[kidx(q, j - 1, d - 2)] * (ne ? 1. : nl_l_loop(&q, i, j - 1, i + 1, j - 2));
    vP *= LC.k0sq;
  }
  double vB = 0.;
  if (gB) {
    const unsigned* ri = q.lf + i * q.mw;
    const unsigned* rj = c.lfr + j * q.mw;
    for (int u0 = 0; u0 <= d; u0 += WARP_N) {
      int u = u0 + lane;
      if (u <= d && row_bit(ri, u) && row_bit(rj, d - u)) vB += t.o1[kidx(q, i, u)] * t.o2[kidx(q, j, d - u)];
    }
    vB = w_sum(vB);
  }
  double v2 = 0., v1 = 0.;
  if (gB) {
    if (ok_B(q, i, d - 1)) v2 += t.o2[kidx(q, j - 1, d - 1)] * LC.k0;
    if (gP) v2 += vP * (ne ? 1. : nl_l_ext(&q, i, j - 1, 0) * el.mlintern);
    v1 = v2 + vB;
  }
  double vM = 0.;
  if (gM) {
    if (ok_M(q, i + 1, d - 1)) vM += t.M[kidx(q, i + 1, d - 1)] * LC.k0;
    if (gB) vM += vB;
  }
  double vE = 0.;
  if (gE) {
    const int C = c.Ceff;
    const int lo = d - C > 0 ? d - C : 0;
    double acc = 0., accg = 0.;
    for (int u10 = 0; u10 <= C; u10 += WARP_N) {
      int u1 = u10 + lane, k = i + u1;
      unsigned m = 0u;
      if (u1 <= C && d - u1 >= lo) {
        m = win_bits(q.bp + k * q.mw, q.mw, lo, d - u1 - lo + 1);
        if (u1 == 0) m &= ~(1u << (d - lo));
      }
      // generic candidates: u1 >= 3 and u2 = d-u1-dd >= 3  <=>  bit index b = dd-lo <= d-u1-3-lo
      unsigned gen = 0u;
      if (!ne && u1 >= 3) {
        int top = d - u1 - 3 - lo;
        if (top >= 0) gen = m & (top >= 31 ? 0xFFFFFFFFu : ((2u << top) - 1u));
      }
      unsigned sp = m & ~gen;
      while (gen) {
        int b = w_ffs(gen) - 1;
        gen &= gen - 1;
        int dd = lo + b, l = k + dd, u2 = d - u1 - dd;
        accg += t.Pm[kidx(q, l, dd)] * ld_ro(t.G + u1 * 32 + u2);
      }
      int ns = 0;
      k0_specials_push(sbuf, ns, sp, u1);
      w_sync();
      if (ns > 128) ns = 128;  // cannot happen: at most 3*31 + 28*3 specials
      for (int z = lane; z < ns; z += WARP_N) {
        int e = sbuf[z], su1 = e >> 8, b = e & 255;
        int sk = i + su1, dd = lo + b, l = sk + dd, u2 = d - su1 - dd;
        acc += t.P[kidx(q, l, dd)] * c.k0pow[su1 + u2] * (ne ? 1. : nl_l_loop(&q, i - 1, j, sk, l - 1));
      }
      w_sync();
    }
    if (!ne) {
      int type = bp_type(q.x[i - 1], q.x[j]);
      accg *= ld_ro(el.mismatch_i + (type * 5 + q.x[i]) * 5 + q.x[j - 1]);
    }
    acc += accg;
    vE = w_sum(acc);
    if (gM) vE += vM * (ne ? 1. : nl_l_ext(&q, j, i - 1, 0) * (el.mlclosing * el.mlintern));
    if (d >= 1) vE += c.k0pow[d] * (ne ? 1. : nl_l_hairpin(&q, i - 1, j));
  }
  if (lane == 0) {
    if (gP) {
      t.P[kidx(q, j, d)] = vP;
      // inner-pair side mismatch of a generic interior loop: pair (i, j-1), neighbours x[j] and x[i-1]
      double mmf = 0.;
      if (i >= 1 && j < q.L) mmf = ld_ro(el.mismatch_i + (bp_type(q.x[j - 1], q.x[i]) * 5 + q.x[j]) * 5 + q.x[i - 1]);
      t.Pm[kidx(q, j, d)] = vP * mmf;
    }
    if (gB) { t.o1[kidx(q, i, d)] = v1; t.o2[kidx(q, j, d)] = v2; }
    if (gM) t.M[kidx(q, i, d)] = vM;
    if (gE) t.E[kidx(q, i, d)] = vE;
  }
}

// exterior row, one warp: O(j) = sum_i O(i) P(i,j) ext(i,j-1) + O(j-1)
RDEV void k0_inside_ext(const LinCtx& c, const K0Tabs& t) {
  const SeqView& q = c.q;
  const DevEnergy& el = LC.el;
  const int L = q.L, lane = lane_id();
  const bool ne = el.no_ene != 0;
  if (lane == 0) { t.O[0] = 1.; t.eO[0] = 0.; }
  w_sync();
  for (int j = 1; j <= L; ++j) {
    const unsigned* rj = c.bpr + j * q.mw;
    int dmax = q.W < j ? q.W : j;
    const int eref = (int)t.eO[j - 1];
    double acc = 0.;
    for (int u0 = 0; u0 <= dmax; u0 += WARP_N) {
      int u = u0 + lane;
      if (u <= dmax && row_bit(rj, u))
        acc += ldexp(t.O[j - u], (int)t.eO[j - u] - eref) * t.P[kidx(q, j, u)] * (ne ? 1. : nl_l_ext(&q, j - u, j - 1, 1));
    }
    acc = w_sum(acc);
    if (lane == 0) {
      double v = acc + t.O[j - 1] * LC.k0;
      int k = renorm_shift(v);
      t.O[j] = k ? ldexp(v, -k) : v;
      t.eO[j] = (double)(eref + k);
    }
    w_sync();
  }
}
RDEV void k0_outside_ext(const LinCtx& c, const K0Tabs& t, double rootw) {
  const SeqView& q = c.q;
  const DevEnergy& el = LC.el;
  const int L = q.L, lane = lane_id();
  const bool ne = el.no_ene != 0;
  // rootw = 1 / mantissa of O(L); the exponent of 1/Z^ is -eO(L)
  if (lane == 0) { t.bO[L] = rootw; t.fO[L] = -t.eO[L]; }
  w_sync();
  for (int i = L - 1; i >= 0; --i) {
    const unsigned* ri = q.bp + i * q.mw;
    int dmax = q.W < L - i ? q.W : L - i;
    const int fref = (int)t.fO[i + 1];
    double acc = 0.;
    for (int u0 = 0; u0 <= dmax; u0 += WARP_N) {
      int u = u0 + lane;
      if (u <= dmax && row_bit(ri, u))
        acc += ldexp(t.bO[i + u], (int)t.fO[i + u] - fref) * t.P[kidx(q, i + u, u)] * (ne ? 1. : nl_l_ext(&q, i, i + u - 1, 1));
    }
    acc = w_sum(acc);
    if (lane == 0) {
      double v = acc + t.bO[i + 1] * LC.k0;
      int k = renorm_shift(v);
      t.bO[i] = k ? ldexp(v, -k) : v;
      t.fO[i] = (double)(fref + k);
    }
    w_sync();
  }
}

RDEV void k0_outside_cell(const LinCtx& c, const K0Tabs& t, int i, int d, int* sbuf) {
  const SeqView& q = c.q;
  const DevEnergy& el = LC.el;
  const int j = i + d, lane = lane_id(), L = q.L, W = q.W;
  const bool ne = el.no_ene != 0;
  const bool gP = ok_P(q, i, d), gB = ok_B(q, i, d), gM = ok_M(q, i, d), gE = ok_E(q, i, d);
  if (!(gP || gB || gM || gE)) return;
  double bE = 0.;
  if (gE) bE = t.bP[kidx(q, i - 1, d + 2)] * LC.k0sq;
  double bM = 0.;
  if (gM) {
    if (gE) bM += bE * (ne ? 1. : nl_l_ext(&q, j, i - 1, 0) * (el.mlclosing * el.mlintern));
    if (ok_M(q, i - 1, d + 1)) bM += t.bM[kidx(q, i - 1, d + 1)] * LC.k0;
  }
  double b1 = 0., bB = 0., b2 = 0.;
  if (gB) {
    {  // this cell as the left child 1(i,k=j) of B(i,j')
      const unsigned* ri = q.lf + i * q.mw;
      const unsigned* rk = q.lf + j * q.mw;
      int dmax = W < L - i ? W : L - i;
      for (int d20 = d; d20 <= dmax; d20 += WARP_N) {
        int d2 = d20 + lane;
        if (d2 <= dmax && row_bit(ri, d2) && row_bit(rk, d2 - d)) b1 += t.bBl[kidx(q, i, d2)] * t.o2[kidx(q, i + d2, d2 - d)];
      }
      b1 = w_sum(b1);
    }
    bB = b1 + (gM ? bM : 0.);
    {  // this cell as the right child 2(k=i,j) of B(i',j)
      const unsigned* rj = c.lfr + j * q.mw;
      int dmax = W < j ? W : j;
      for (int d20 = d; d20 <= dmax; d20 += WARP_N) {
        int d2 = d20 + lane, i2 = j - d2;
        if (d2 <= dmax && row_bit(rj, d2) && row_bit(q.lf + i2 * q.mw, d2 - d))
          b2 += t.bBr[kidx(q, j, d2)] * t.o1[kidx(q, i2, d2 - d)];
      }
      b2 = w_sum(b2) + b1;
    }
    if (ok_B(q, i, d + 1)) b2 += t.b2[kidx(q, i, d + 1)] * LC.k0;
  }
  double bP = 0.;
  if (gP) {
    if (gB) bP += b2 * (ne ? 1. : nl_l_ext(&q, i, j - 1, 0) * el.mlintern);
    if (ok_P(q, i - 1, d + 2)) bP += t.bP[kidx(q, i - 1, d + 2)] * LC.k0sq * (ne ? 1. : nl_l_loop(&q, i - 1, j, i, j - 1));
    bP += ldexp(t.O[i] * t.bO[j], (int)(t.eO[i] + t.fO[j])) * (ne ? 1. : nl_l_ext(&q, i, j - 1, 1));
    // enclosing pairs: this cell is the inner pair (k=i,l=j) of E(i',j')
    const int C = c.Ceff;
    const int hi = W < d + c.Csum + 2 ? W : d + c.Csum + 2;
    double acc = 0., accg = 0.;
    for (int u10 = 0; u10 <= C; u10 += WARP_N) {
      int u1 = u10 + lane, i2 = i - u1, lo = d + u1 + 2;
      unsigned m = 0u;
      if (u1 <= C && i2 >= 1 && hi >= lo) {
        m = win_bits(q.bp + (i2 - 1) * q.mw, q.mw, lo, hi - lo + 1);
        if (u1 == 0) m &= ~1u;
      }
      unsigned gen = (!ne && u1 >= 3) ? (m & ~7u) : 0u;   // bit index = u2
      unsigned sp = m & ~gen;
      while (gen) {
        int u2 = w_ffs(gen) - 1;
        gen &= gen - 1;
        accg += t.bEm[kidx(q, i2, d + u1 + u2)] * ld_ro(t.G + u1 * 32 + u2);
      }
      int ns = 0;
      k0_specials_push(sbuf, ns, sp, u1);
      w_sync();
      if (ns > 128) ns = 128;
      for (int z = lane; z < ns; z += WARP_N) {
        int e = sbuf[z], su1 = e >> 8, u2 = e & 255;
        int si2 = i - su1, j2 = j + u2;
        acc += t.bE[kidx(q, si2, d + su1 + u2)] * c.k0pow[su1 + u2] * (ne ? 1. : nl_l_loop(&q, si2 - 1, j2, i, j - 1));
      }
      w_sync();
    }
    if (!ne && i >= 1 && j < L)
      accg *= ld_ro(el.mismatch_i + (bp_type(q.x[j - 1], q.x[i]) * 5 + q.x[j]) * 5 + q.x[i - 1]);
    else accg = 0.;
    acc += accg;
    bP += w_sum(acc);
  }
  if (lane == 0) {
    if (gP) t.bP[kidx(q, i, d)] = bP;
    if (gE) {
      t.bE[kidx(q, i, d)] = bE;
      // closing-pair side mismatch of a generic interior loop: pair (i-1, j), neighbours x[i] and x[j-1]
      t.bEm[kidx(q, i, d)] = bE * ld_ro(el.mismatch_i + (bp_type(q.x[i - 1], q.x[j]) * 5 + q.x[i]) * 5 + q.x[j - 1]);
    }
    if (gM) t.bM[kidx(q, i, d)] = bM;
    if (gB) { t.bBl[kidx(q, i, d)] = bB; t.bBr[kidx(q, j, d)] = bB; t.b2[kidx(q, i, d)] = b2; }
  }
}

#if LIN_SPLIT_TMA && !defined(RELEM_HOST_EMU)
// ---- bulk-copy (TMA) primitives, sm_90+ PTX
RDEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
RDEV void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
RDEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
RDEV void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
RDEV void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#endif

// ================================================================================= coupled passes
struct CTabs {
  double *aP, *aE, *aM, *a1, *a2, *aLl, *aLr, *aO;          // aP, a2, aLr by right end; the rest by left end
  double *bP, *bEl, *bEr, *bM, *bBl, *bBr, *b2, *bL, *bO;   // bEr, bBr by right end; channel c at + c*bch (bO: + c*boch)
  double *eO, *fO;   // [L+1] power-of-two exponents of the exterior rows aO / bO (all channels share fO)
  unsigned bch, boch;
};

struct WarpLin {
  double* curA;   // [S] staging of the current cell (inside: B; exterior rows)
  double* curB;   // [NCH][S] outside values of the state type being gathered
  double* partA;  // [NCH][n_max] per-entry partial sums
  double* partT;  // [NCH][n_max] per-entry partial sums weighted by the transition energy
  int *bi, *bj;   // candidate buffer
  double *bt, *bf0, *bf1;
  int* kbuf;      // [W+2] compacted split points
  double* cntR;   // [NCH][n_right*5] emission posterior sums per (entry, base), private to the warp
  double* cntL;   // [NCH][n_left*5]
  double* pcnt;   // pair-emission posterior sums [n_pair*25] of channel 0 in the slot header (global, RED)
  unsigned pstride;  // channel stride of pcnt
  int n_max;
  double* stage;             // LIN_SPLIT_TMA: [LIN_TMA_STAGES][2][S] staged operand vectors of the split gather
  unsigned long long* bar;   // its mbarrier
  unsigned phase;            // parity of the next completion
};
// longest list the OUTSIDE phases index their per-entry sums with: in scatter mode the quads (by far the longest list
// of a large automaton: 1 820 against 455 splits at S = 91) are walked without per-entry sums, so their count does not
// size the per-warp shared memory any more (184 KB -> 88 KB per CTA at S = 91: two CTAs per SM instead of one)
RHD int lin_outside_nmax(const LinHMM& h) {
#if LIN_SCATTER_ILOOP
  int m = h.S;
  if (h.n_right > m) m = h.n_right;
  if (h.n_left > m) m = h.n_left;
  if (h.n_pair > m) m = h.n_pair;
  if (h.n_split > m) m = h.n_split;
  return m;
#else
  return h.n_max;
#endif
}
// inside = true: only what the inside pass needs (no outside staging, no counts)
// staged = true: room for the bulk-copy staging buffer of the split gather (inside B only)
// n_max: longest list an outside / inside phase keeps per-entry sums for (lin_outside_nmax for the outside phases)
RHD int warp_lin_bytes(int S, int Wmax, int nch, int n_max, int n_right, int n_left, bool inside, bool staged = false) {
  // scatter mode keeps no energy-weighted per-entry sums (partT: only the enclosing-loop gather of outside P used them)
  int n = inside ? (S + n_max + 3 * LIN_CAP) * 8
                 : (S + nch * S + (LIN_SCATTER_ILOOP ? 1 : 2) * nch * n_max + 3 * LIN_CAP + nch * 5 * (n_right + n_left)) * 8;
  n += (2 * LIN_CAP + Wmax + 4) * 4;
  n = (n + 15) & ~15;
#if LIN_SPLIT_TMA
  if (inside && staged) n += LIN_TMA_STAGES * 2 * S * 8 + 16;   // staging buffer (16-byte aligned: n is) + mbarrier
#else
  (void)staged;
#endif
  return n;
}
RDEV WarpLin warp_lin_carve(unsigned char* base, int S, int Wmax, int nch, int n_max, int n_right, int n_left, bool inside,
                            bool staged = false) {
  WarpLin w;
  double* p = (double*)base;
  w.curA = p; p += S;
  w.curB = p; p += inside ? 0 : nch * S;
  w.partA = p; p += inside ? n_max : nch * n_max;
  w.partT = p; p += (inside || LIN_SCATTER_ILOOP) ? 0 : nch * n_max;
  w.bt = p; p += LIN_CAP;
  w.bf0 = p; p += LIN_CAP;
  w.bf1 = p; p += LIN_CAP;
  w.cntR = p; p += inside ? 0 : nch * 5 * n_right;
  w.cntL = p; p += inside ? 0 : nch * 5 * n_left;
  int* ip = (int*)p;
  w.bi = ip; ip += LIN_CAP;
  w.bj = ip; ip += LIN_CAP;
  w.kbuf = ip;
  w.pcnt = nullptr;
  w.n_max = n_max;
  w.stage = nullptr; w.bar = nullptr; w.phase = 0u;
#if LIN_SPLIT_TMA
  if (inside && staged) {
    int n = (S + n_max + 3 * LIN_CAP) * 8 + (2 * LIN_CAP + Wmax + 4) * 4;
    n = (n + 15) & ~15;
    w.stage = (double*)(base + n);
    w.bar = (unsigned long long*)(base + n + LIN_TMA_STAGES * 2 * S * 8);
  }
#else
  (void)Wmax; (void)staged;
#endif
  return w;
}

RDEV double seg_sum(const double* part, const int* off, int s) {
  // segments are short (mostly 0..3 entries): straight-line code for the first three, a loop for the rest
  const int a = ld_ro(off + s), n = ld_ro(off + s + 1) - a;
  double v = 0.;
  if (n > 0) v = part[a];
  if (n > 1) v += part[a + 1];
  if (n > 2) v += part[a + 2];
  for (int k = 3; k < n; ++k) v += part[a + k];
  return v;
}

// all lanes call; lanes with ok append (ia, ib, tsc) and the two Boltzmann factors exp(lambda_slot * tsc)
// Structural candidates are collected in two stages: the mask scans only append (ia, ib) ids (ballot compaction);
// once a batch is complete its energies are evaluated with ONE candidate per lane, so the ~300-instruction case
// analysis + two exponentials run at full lane occupancy however sparse the scan was.  Candidates with a forbidden
// energy stay in the batch with weight 0.
RDEV void batch_add(WarpLin& w, int& n, bool ok, int ia, int ib) {
  unsigned bal = w_ballot(ok);
  if (ok) {
    int pos = n + w_popc(bal & lanemask_lt());
    w.bi[pos] = ia; w.bj[pos] = ib;
  }
  n += w_popc(bal);
}
// energy(ia, ib) -> tsc (log-Boltzmann weight of the structural transition)
template <class E> RDEV void batch_eval(WarpLin& w, int n, E energy) {
  w_sync();
  const bool ne = LC.en.no_ene != 0;
  for (int z = lane_id(); z < n; z += WARP_N) {
    double tsc = 0., f0 = 1., f1 = 1.;
    if (!ne) {
      tsc = energy(w.bi[z], w.bj[z]);
      if (tsc > NINF) { F2 ff = boltz2(tsc); f0 = ff.f0; f1 = ff.f1; }
      else { tsc = 0.; f0 = 0.; f1 = 0.; }
    }
    w.bt[z] = tsc; w.bf0[z] = f0; w.bf1[z] = f1;
  }
  w_sync();
}
#define LIN_ROOM(n, energy, flush)                    \
  if ((n) > LIN_CAP - WARP_N) {                       \
    batch_eval(w, n, energy);                         \
    flush(n);                                         \
    (n) = 0;                                          \
    w_sync();                                         \
  }
#define LIN_DONE(n, energy, flush)                    \
  if (n) {                                            \
    batch_eval(w, n, energy);                         \
    flush(n);                                         \
    w_sync();                                         \
  }

// outside-pass limits of a sequence with band W (see LinCtx::Csum)
RDEV void lin_outside_limits(int W, int max_iloop, bool no_ene, int Ceff, int& Csum, int& Cfl) {
  const bool binding = !no_ene && max_iloop < 30 && max_iloop < W - 7;
  Csum = binding ? 30 : Ceff;
  Cfl = binding ? (W < 30 ? W : 30) : Ceff;
}

// inner pairs (k,l) of E(i,j): u1 = k-i <= C, u2 = j-l, u1+u2 <= csum, not both 0 (energy_model.hpp:413-426; csum = C
// in the inside pass, LinCtx::Csum in the outside pass)
template <class F> RDEV void walk_inner(const LinCtx& c, int i, int d, WarpLin& w, int csum, F flush) {
  const SeqView& q = c.q;
  const int j = i + d, C = c.Ceff, lane = lane_id();
  const int lo = d - csum > 0 ? d - csum : 0;
  auto energy = [&](int k, int l) { return nl_e_loop(&q, i - 1, j, k, l - 1); };
  int n = 0;
  for (int u10 = 0; u10 <= C; u10 += WARP_N) {
    int u1 = u10 + lane, k = i + u1;
    unsigned m = 0u;
    if (u1 <= C && d - u1 >= lo) {
      m = win_bits(q.bp + k * q.mw, q.mw, lo, d - u1 - lo + 1);
      if (u1 == 0) m &= ~(1u << (d - lo));
    }
    while (w_any(m != 0u)) {
      LIN_ROOM(n, energy, flush)
      bool has = m != 0u;
      int b = 0;
      if (has) { b = w_ffs(m) - 1; m &= m - 1; }
      batch_add(w, n, has, k, k + lo + b);
    }
  }
  LIN_DONE(n, energy, flush)
}
// enclosing pairs: E(i',j') that have (k=i,l=j) as inner pair
template <class F> RDEV void walk_outer(const LinCtx& c, int i, int d, WarpLin& w, F flush) {
  const SeqView& q = c.q;
  const int j = i + d, C = c.Ceff, lane = lane_id(), W = q.W;
  const int hi = W < d + C + 2 ? W : d + C + 2;
  auto energy = [&](int i2, int j2) { return nl_e_loop(&q, i2 - 1, j2, i, j - 1); };
  int n = 0;
  for (int u10 = 0; u10 <= C; u10 += WARP_N) {
    int u1 = u10 + lane, i2 = i - u1, lo = d + u1 + 2;
    unsigned m = 0u;
    if (u1 <= C && i2 >= 1 && hi >= lo) {
      m = win_bits(q.bp + (i2 - 1) * q.mw, q.mw, lo, hi - lo + 1);
      if (u1 == 0) m &= ~1u;
    }
    while (w_any(m != 0u)) {
      LIN_ROOM(n, energy, flush)
      bool has = m != 0u;
      int u2 = 0;
      if (has) { u2 = w_ffs(m) - 1; m &= m - 1; }
      batch_add(w, n, has, i2, j + u2);
    }
  }
  LIN_DONE(n, energy, flush)
}
// cell (i,k) as the left unpaired flank L(i,k) of E(i,j') with inner pair (k,l); pushes (l, j')
template <class F> RDEV void walk_left_flank(const LinCtx& c, int i, int d, WarpLin& w, F flush) {
  const SeqView& q = c.q;
  const int k = i + d, C = c.Ceff, lane = lane_id(), W = q.W;
  const unsigned* rk = q.bp + k * q.mw;
  const unsigned* ri = q.bp + (i - 1) * q.mw;
  auto energy = [&](int l, int j2) { return nl_e_loop(&q, i - 1, j2, k, l - 1); };
  int n = 0;
  for (int dd0 = 0; dd0 <= W; dd0 += WARP_N) {
    int dd = dd0 + lane, l = k + dd;
    unsigned m = 0u;
    int lo = d + dd + 2;
    if (dd <= W && row_bit(rk, dd)) {
      int hi = W < lo + C - d ? W : lo + C - d;
      if (hi >= lo) m = win_bits(ri, q.mw, lo, hi - lo + 1);
    }
    while (w_any(m != 0u)) {
      LIN_ROOM(n, energy, flush)
      bool has = m != 0u;
      int u2 = 0;
      if (has) { u2 = w_ffs(m) - 1; m &= m - 1; }
      batch_add(w, n, has, l, l + u2);
    }
  }
  LIN_DONE(n, energy, flush)
}
// cell (l,j) as the right unpaired flank L(l,j) of E(i',j) with inner pair (k,l); pushes (k, i')
template <class F> RDEV void walk_right_flank(const LinCtx& c, int l, int d, WarpLin& w, F flush) {
  const SeqView& q = c.q;
  const int j = l + d, C = c.Ceff, lane = lane_id(), W = q.W;
  const unsigned* rl = c.bpr + l * q.mw;
  const unsigned* rj = c.bpr + (j + 1) * q.mw;
  auto energy = [&](int k, int i2) { return nl_e_loop(&q, i2 - 1, j, k, l - 1); };
  int n = 0;
  for (int dd0 = 0; dd0 <= W; dd0 += WARP_N) {
    int dd = dd0 + lane, k = l - dd;
    unsigned m = 0u;
    int lo = d + dd + 2;
    if (dd <= W && k >= 0 && row_bit(rl, dd)) {
      int hi = W < lo + C - d ? W : lo + C - d;
      if (hi >= lo) m = win_bits(rj, q.mw, lo, hi - lo + 1);
    }
    while (w_any(m != 0u)) {
      LIN_ROOM(n, energy, flush)
      bool has = m != 0u;
      int u1 = 0;
      if (has) { u1 = w_ffs(m) - 1; m &= m - 1; }
      batch_add(w, n, has, k, k - u1);
    }
  }
  LIN_DONE(n, energy, flush)
}

// ------------------------------------------------------------------------------------------------- inside
// The cell update is split into four PHASES that a warp runs one after the other over all of its cells of a
// diagonal (phase-major order): the instruction working set at any time is one phase (a few hundred instructions)
// instead of the whole cell update, which matters because the SM's instruction cache holds only ~2K instructions.
// Values that flow between phases of the same cell (P -> 2, B -> M, {M,L} -> E) go through the tables themselves
// (same warp, L1/L2 hits).
//
// ---- phase L: L(i,j,s) <- L(i,j-1,s1) emitR   (every cell)
RDEV void lin_in_L(const LinCtx& c, const CTabs& t, int i, int d, WarpLin& w) {
  const LinHMM& h = LC.h;
  const LinParams& p = LC.p;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, j = i + d, lane = lane_id();
  const unsigned il = cidx(q, i, d), ir = cidx(q, j, d);
  if (d == 0) {
    for (int s = lane; s < S; s += WARP_N) {
      double v = ld_ro(h.slot + s) ? 0. : 1.;
      t.aLl[il + s] = v;
      t.aLr[ir + s] = v;
    }
    return;
  }
  double* part = w.partA;
  const int xr = q.x[j - 1];
  const double wsr = c.wsf[j - 1];
  const bool atR = (j - 1 == c.ys);
  const double* src = t.aLl + cidx(q, i, d - 1);
  for (int a = lane; a < h.n_right; a += WARP_N) {
    int fl = ld_ro(h.r_flag + a);
    double v = 0.;
    if (fl & 2) {
      v = src[ld_ro(h.r_src + a)] * ld_ro(p.r_w + a * 5 + xr);
      if (fl & 1) v *= wsr;
      if (atR && !(fl & LIN_F_START)) v = 0.;
    }
    part[a] = v;
  }
  w_sync();
  for (int s = lane; s < S; s += WARP_N) {
    double v = seg_sum(part, h.r_off, s);
    t.aLl[il + s] = v;
    t.aLr[ir + s] = v;
  }
  w_sync();
}

// ---- phase P: P(i,j,s) <- E(i+1,j-1,s1) | P(i+1,j-1,s1)   (cells with an allowed pair)
RDEV void lin_in_P(const LinCtx& c, const CTabs& t, int i, int d, WarpLin& w) {
  const LinHMM& h = LC.h;
  const LinParams& p = LC.p;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, j = i + d, lane = lane_id();
  double* part = w.partA;
  const bool ne = LC.en.no_ene != 0;
  const int xl = q.x[i], xr = q.x[j - 1];
  const double wsl = c.wsf[i], wsr = c.wsf[j - 1];
  const bool cE = ok_E(q, i + 1, d - 2), cP = ok_P(q, i + 1, d - 2);
  bool cPP = cP;
  double f0 = 1., f1 = 1.;
  if (cP && !ne) {
    double tsc = nl_e_loop(&c.q, i, j - 1, i + 1, j - 2);
    cPP = tsc > NINF;
    if (cPP) { F2 ff = boltz2(tsc); f0 = ff.f0; f1 = ff.f1; }
  }
  const double* srcE = t.aE + cidx(q, i + 1, d - 2);
  const double* srcP = t.aP + cidx(q, j - 1, d - 2);
  for (int a = lane; a < h.n_pair; a += WARP_N) {
    double v = 0.;
    if (cE || cPP) {
      int fl = ld_ro(h.p_flag + a), s1 = ld_ro(h.p_src + a);
      double wt = ld_ro(p.p_w + a * 25 + xl * 5 + xr);
      if (fl & 1) wt *= wsl;
      if (fl & 2) wt *= wsr;
      if ((i == c.ys && !(fl & LIN_F_START)) || (j - 1 == c.ys && !(fl & LIN_F_RSTART))) wt = 0.;
      if (cE) v += srcE[s1] * wt;
      if (cPP) v += srcP[s1] * wt * (ld_ro(h.slot + ld_ro(h.p_tgt + a)) ? f1 : f0);
    }
    part[a] = v;
  }
  w_sync();
  const unsigned ir = cidx(q, j, d);
  for (int s = lane; s < S; s += WARP_N) t.aP[ir + s] = seg_sum(part, h.p_off, s);
  w_sync();
}

// ---- phase B: B <- 1 2;  2 <- 2 emitR | P;  1 <- 2 | B;  M <- M emitL | B   (cells with gB or gM)
RDEV void lin_in_B(const LinCtx& c, const CTabs& t, int i, int d, bool gP, bool gB, bool gM, WarpLin& w) {
  const LinHMM& h = LC.h;
  const LinParams& p = LC.p;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, j = i + d, lane = lane_id();
  double* cur = w.curA;   // [0..S) B of this cell
  double* part = w.partA;
  const bool ne = LC.en.no_ene != 0;
  const unsigned il = cidx(q, i, d), ir = cidx(q, j, d);
  if (gB) {
    int nk = 0;
    {
      const unsigned* ri = q.lf + i * q.mw;
      const unsigned* rj = c.lfr + j * q.mw;
      for (int u0 = 0; u0 <= d; u0 += WARP_N) {
        int u = u0 + lane;
        bool ok = u <= d && row_bit(ri, u) && row_bit(rj, d - u);
        unsigned bal = w_ballot(ok);
        if (ok) w.kbuf[nk + w_popc(bal & lanemask_lt())] = u;
        nk += w_popc(bal);
      }
      w_sync();
    }
    const double* r1 = t.a1 + cidx(q, i, 0);
    const double* r2 = t.a2 + cidx(q, j, 0) + (unsigned)d * S;
#if LIN_SPLIT_TMA && !defined(RELEM_HOST_EMU)
    if ((S & 1) == 0 && w.stage) {
      // staged: batches of up to LIN_TMA_STAGES split points; lane t of a batch issues the two bulk copies of its split
      // point (1(i,u,.) and 2(u,j,.), S doubles each), lane 0 arms the barrier with the batch's byte count
      const unsigned vb = (unsigned)S * 8u;
      for (int a = lane; a < h.n_split; a += WARP_N) part[a] = 0.;
      for (int t0 = 0; t0 < nk; t0 += LIN_TMA_STAGES) {
        const int nb = nk - t0 < LIN_TMA_STAGES ? nk - t0 : LIN_TMA_STAGES;
        w_sync();   // the previous batch has been consumed
        if (lane == 0) mbar_expect_tx(w.bar, 2u * vb * (unsigned)nb);
        w_sync();
        if (lane < nb) {
          const int u = w.kbuf[t0 + lane] * S;
          bulk_g2s(w.stage + (2 * lane) * S, r1 + u, vb, w.bar);
          bulk_g2s(w.stage + (2 * lane + 1) * S, r2 - u, vb, w.bar);
        }
        mbar_wait(w.bar, w.phase);
        w.phase ^= 1u;
        for (int a = lane; a < h.n_split; a += WARP_N) {
          const int sl = ld_ro(h.sp_l + a), sr = ld_ro(h.sp_r + a);
          double v0 = 0., v1 = 0.;
          int tt = 0;
          for (; tt + 1 < nb; tt += 2) {
            v0 += w.stage[(2 * tt) * S + sl] * w.stage[(2 * tt + 1) * S + sr];
            v1 += w.stage[(2 * tt + 2) * S + sl] * w.stage[(2 * tt + 3) * S + sr];
          }
          if (tt < nb) v0 += w.stage[(2 * tt) * S + sl] * w.stage[(2 * tt + 1) * S + sr];
          part[a] += v0 + v1;
        }
      }
    } else
#endif
    for (int a = lane; a < h.n_split; a += WARP_N) {
      const double* p1 = r1 + ld_ro(h.sp_l + a);
      const double* p2 = r2 + ld_ro(h.sp_r + a);
      double v0 = 0., v1 = 0.;
      int tt = 0;
#if LIN_SPLIT_UNROLL >= 4
      double v2 = 0., v3 = 0.;
      for (; tt + 3 < nk; tt += 4) {   // four split points in flight
        int u0 = w.kbuf[tt] * S, u1 = w.kbuf[tt + 1] * S, u2 = w.kbuf[tt + 2] * S, u3 = w.kbuf[tt + 3] * S;
        double a0 = p1[u0], a1 = p1[u1], a2 = p1[u2], a3 = p1[u3];
        double b0 = p2[-u0], b1 = p2[-u1], b2 = p2[-u2], b3 = p2[-u3];
        v0 += a0 * b0; v1 += a1 * b1; v2 += a2 * b2; v3 += a3 * b3;
      }
      v0 += v2; v1 += v3;
#endif
      for (; tt + 1 < nk; tt += 2) {
        int u0 = w.kbuf[tt] * S, u1 = w.kbuf[tt + 1] * S;
        v0 += p1[u0] * p2[-u0];
        v1 += p1[u1] * p2[-u1];
      }
      if (tt < nk) { int u0 = w.kbuf[tt] * S; v0 += p1[u0] * p2[-u0]; }
      part[a] = v0 + v1;
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N) cur[s] = seg_sum(part, h.sp_off, s);
    w_sync();
    const int xr = q.x[j - 1];
    const double wsr = c.wsf[j - 1];
    const bool ok2 = ok_B(q, i, d - 1);
    const double* src2 = t.a2 + cidx(q, j - 1, d - 1);
    for (int a = lane; a < h.n_right; a += WARP_N) {
      double v = 0.;
      if (ok2) {
        int fl = ld_ro(h.r_flag + a);
        v = src2[ld_ro(h.r_src + a)] * ld_ro(p.r_w + a * 5 + xr);
        if (fl & 1) v *= wsr;
        if (j - 1 == c.ys && !(fl & LIN_F_START)) v = 0.;
      }
      part[a] = v;
    }
    bool c2P = gP;
    double f0 = 1., f1 = 1.;
    if (gP && !ne) {
      double tsc = nl_e_ext(&c.q, i, j - 1, 0) + LC.en.mlintern;
      c2P = tsc > NINF;
      if (c2P) { F2 ff = boltz2(tsc); f0 = ff.f0; f1 = ff.f1; }
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N) {
      double x = seg_sum(part, h.r_off, s);
      if (c2P) x += t.aP[ir + s] * (ld_ro(h.slot + s) ? f1 : f0);
      t.a2[ir + s] = x;
      t.a1[il + s] = x + cur[s];
    }
    w_sync();
  }
  if (gM) {
    const int xl = q.x[i];
    const double wsl = c.wsf[i];
    const bool okM = ok_M(q, i + 1, d - 1);
    const double* srcM = t.aM + cidx(q, i + 1, d - 1);
    for (int a = lane; a < h.n_left; a += WARP_N) {
      double v = 0.;
      if (okM) {
        int fl = ld_ro(h.l_flag + a);
        v = srcM[ld_ro(h.l_src + a)] * ld_ro(p.l_w + a * 5 + xl);
        if (fl & 1) v *= wsl;
        if (i == c.ys && !(fl & LIN_F_START)) v = 0.;
      }
      part[a] = v;
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N) {
      double x = seg_sum(part, h.l_off, s);
      if (gB) x += cur[s];
      t.aM[il + s] = x;
    }
    w_sync();
  }
}

// ---- phase E: E(i,j,s) <- M(i,j,s) | L(i,j,s) hairpin | P(k,l,s1) L(i,k,s2) L(l,j,s3)   (cells enclosed by a pair)
RDEV void lin_in_E(const LinCtx& c, const CTabs& t, int i, int d, bool gM, WarpLin& w) {
  const LinHMM& h = LC.h;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, j = i + d, lane = lane_id();
  double* part = w.partA;
  const bool ne = LC.en.no_ene != 0;
  const unsigned il = cidx(q, i, d);
  for (int a = lane; a < h.n_quad; a += WARP_N) part[a] = 0.;
  w_sync();
  if (h.n_quad > 0) {
    const double* rL = t.aLl + cidx(q, i, 0);
    const double* rR = t.aLr + cidx(q, j, 0);
    walk_inner(c, i, d, w, c.Ceff, [&](int n) {
      for (int a = lane; a < h.n_quad; a += WARP_N) {
        int s1 = ld_ro(h.q_s1 + a), s2 = ld_ro(h.q_s2 + a), s3 = ld_ro(h.q_s3 + a);
        const double* bf = ld_ro(h.slot + ld_ro(h.q_tgt + a)) ? w.bf1 : w.bf0;
        double v = part[a];
        LIN_PP_PRAGMA
        for (int pp = 0; pp < n; ++pp) {
          int k = w.bi[pp], l = w.bj[pp];
          double a0 = t.aP[cidx(q, l, l - k) + s1];
          v += a0 * rL[(unsigned)(k - i) * S + s2] * rR[(unsigned)(j - l) * S + s3] * bf[pp];
        }
        part[a] = v;
      }
    });
  }
  bool cM = gM, cH = true;
  double m0 = 1., m1 = 1., h0 = 1., h1 = 1.;
  if (!ne) {
    if (gM) {
      double tM = nl_e_ext(&c.q, j, i - 1, 0) + (LC.en.mlclosing + LC.en.mlintern);
      cM = tM > NINF;
      if (cM) { F2 ff = boltz2(tM); m0 = ff.f0; m1 = ff.f1; }
    }
    double tH = nl_e_hairpin(&c.q, i - 1, j);
    cH = tH > NINF;
    if (cH) { F2 ff = boltz2(tH); h0 = ff.f0; h1 = ff.f1; }
  }
  for (int s = lane; s < S; s += WARP_N) {
    double x = seg_sum(part, h.q_off, s);
    int sl = ld_ro(h.slot + s);
    if (cM) x += t.aM[il + s] * (sl ? m1 : m0);
    if (cH && ld_ro(h.is_loop + s)) x += t.aLl[il + s] * (sl ? h1 : h0);
    t.aE[il + s] = x;
  }
  w_sync();
}

// exterior row, one warp: O(j,s) <- O(i,(s.l,h)) P(i,j,(h,s.r)) ext | O(j-1,s1) emitR
RDEV void lin_inside_ext(const LinCtx& c, const CTabs& t, WarpLin& w) {
  const LinHMM& h = LC.h;
  const LinParams& p = LC.p;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, L = q.L, lane = lane_id();
  const bool ne = LC.en.no_ene != 0;
  double* part = w.partA;
  double* cur = w.curA;
  for (int s = lane; s < S; s += WARP_N) t.aO[s] = (s == h.s00) ? 1. : 0.;
  if (lane == 0) t.eO[0] = 0.;
  w_sync();
  for (int j = 1; j <= L; ++j) {
    const int eref = (int)t.eO[j - 1];   // column j is assembled at the scale of column j-1, then rescaled
    for (int a = lane; a < h.n_split; a += WARP_N) part[a] = 0.;
    w_sync();
    auto flush = [&](int n) {
      // bring every candidate's O(i) to the scale of this column: once per candidate, folded into its factors
      for (int z = lane; z < n; z += WARP_N) {
        double sc = ldexp(1., (int)t.eO[w.bi[z]] - eref);
        w.bf0[z] *= sc; w.bf1[z] *= sc;
      }
      w_sync();
      for (int a = lane; a < h.n_split; a += WARP_N) {
        int sl = ld_ro(h.sp_l + a), sr = ld_ro(h.sp_r + a);
        const double* bf = ld_ro(h.slot + ld_ro(h.sp_tgt + a)) ? w.bf1 : w.bf0;
        double v = part[a];
        LIN_PP_PRAGMA
        for (int pp = 0; pp < n; ++pp) {
          int i = w.bi[pp];
          v += t.aO[(unsigned)i * S + sl] * t.aP[cidx(q, j, j - i) + sr] * bf[pp];
        }
        part[a] = v;
      }
    };
    {
      const unsigned* rj = c.bpr + j * q.mw;
      int dmax = q.W < j ? q.W : j, n = 0;
      auto energy = [&](int i, int) { return nl_e_ext(&c.q, i, j - 1, 1); };
      for (int u0 = 0; u0 <= dmax; u0 += WARP_N) {
        LIN_ROOM(n, energy, flush)
        int u = u0 + lane;
        batch_add(w, n, u <= dmax && row_bit(rj, u), j - u, j);
      }
      LIN_DONE(n, energy, flush)
    }
    for (int s = lane; s < S; s += WARP_N) cur[s] = seg_sum(part, h.sp_off, s);
    w_sync();
    const int xr = q.x[j - 1];
    const double wsr = c.wsf[j - 1];
    for (int a = lane; a < h.n_right; a += WARP_N) {
      int fl = ld_ro(h.r_flag + a);
      double v = t.aO[(unsigned)(j - 1) * S + ld_ro(h.r_src + a)] * ld_ro(p.r_w + a * 5 + xr);
      if (fl & 1) v *= wsr;
      if (j - 1 == c.ys && !(fl & LIN_F_START)) v = 0.;
      part[a] = v;
    }
    w_sync();
    double mx = 0.;
    for (int s = lane; s < S; s += WARP_N) {
      double v = cur[s] + seg_sum(part, h.r_off, s);
      cur[s] = v;
      mx = fmax(mx, fabs(v));
    }
    const int k = renorm_shift(w_max(mx));
    for (int s = lane; s < S; s += WARP_N) t.aO[(unsigned)j * S + s] = k ? ldexp(cur[s], -k) : cur[s];
    if (lane == 0) t.eO[j] = (double)(eref + k);
    w_sync();
  }
}

// ------------------------------------------------------------------------------------------------ outside
// eh[c*2+slot]: lambda-gradient sums (sum of tsc * posterior), per lane, reduced by the caller
template <int NCH> struct EhAcc {
  double v[NCH * 2];
  RDEV void add(int c, int slot, double x) {
    v[c * 2] += slot ? 0. : x;
    v[c * 2 + 1] += slot ? x : 0.;
  }
};

// exterior row top-down, one warp.  bO(L,.) must hold the root weights.
template <int NCH, int MODE = 0> RDEV void lin_outside_ext(const LinCtx& c, const CTabs& t, WarpLin& w) {
  const LinHMM& h = LC.h;
  const LinParams& p = LC.p;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, L = q.L, lane = lane_id(), NM = w.n_max;
  const bool ne = LC.en.no_ene != 0;
  for (int i = L - 1; i >= 0; --i) {
    const int fref = (int)t.fO[i + 1];   // row i is assembled at the scale of row i+1, then rescaled
    for (int a = lane; a < h.n_split; a += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) w.partA[ch * NM + a] = 0.;
    w_sync();
    auto flush = [&](int n) {
      for (int z = lane; z < n; z += WARP_N) {
        double sc = ldexp(1., (int)t.fO[w.bi[z]] - fref);
        w.bf0[z] *= sc; w.bf1[z] *= sc;
      }
      w_sync();
      for (int pz = lane; pz < h.n_split; pz += WARP_N) {
        int a = ld_ro(h.spL_ord + pz);
        int s = ld_ro(h.sp_tgt + a), sr = ld_ro(h.sp_r + a);
        const double* bf = ld_ro(h.slot + s) ? w.bf1 : w.bf0;
        double v[NCH];
        for (int ch = 0; ch < NCH; ++ch) v[ch] = w.partA[ch * NM + pz];
        LIN_PP_PRAGMA
        for (int pp = 0; pp < n; ++pp) {
          int j = w.bi[pp];
          double term = t.aP[cidx(q, j, j - i) + sr] * bf[pp];
          for (int ch = 0; ch < NCH; ++ch) v[ch] += t.bO[ch * t.boch + (unsigned)j * S + s] * term;
        }
        for (int ch = 0; ch < NCH; ++ch) w.partA[ch * NM + pz] = v[ch];
      }
    };
    {
      const unsigned* ri = q.bp + i * q.mw;
      int dmax = q.W < L - i ? q.W : L - i, n = 0;
      auto energy = [&](int j, int) { return nl_e_ext(&c.q, i, j - 1, 1); };
      for (int u0 = 0; u0 <= dmax; u0 += WARP_N) {
        LIN_ROOM(n, energy, flush)
        int u = u0 + lane;
        batch_add(w, n, u <= dmax && row_bit(ri, u), i + u, i);
      }
      LIN_DONE(n, energy, flush)
    }
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) w.curB[ch * S + s] = seg_sum(w.partA + ch * NM, h.spL_off, s);
    w_sync();
    // O(i+1,sp) -> O(i,s1) emitting x[i]
    const int xr = q.x[i];
    const double wsr = c.wsf[i];
    for (int pz = lane; pz < h.n_right; pz += WARP_N) {
      int a = ld_ro(h.rT_ord + pz);
      int sp = ld_ro(h.r_tgt + a), ch_s = ld_ro(h.r_src + a), fl = ld_ro(h.r_flag + a);
      double wt = ld_ro(p.r_w + a * 5 + xr);
      if (fl & 1) wt *= wsr;
      if (i == c.ys && !(fl & LIN_F_START)) wt = 0.;
      // posterior = b^O(i+1) w a^O(i): mantissas times 2^(fO(i+1) + eO(i))
      double ac = ldexp(t.aO[(unsigned)i * S + ch_s], fref + (int)t.eO[i]);
      for (int ch = 0; ch < NCH; ++ch) {
        double contrib = t.bO[ch * t.boch + (unsigned)(i + 1) * S + sp] * wt;
        w.partA[ch * NM + pz] = contrib;
        if (MODE != 2 && !p.no_prf) w.cntR[(ch * h.n_right + a) * 5 + xr] += contrib * ac;
        if (MODE != 0 && ch == 0 && contrib != 0.) hook_emit_right<MODE>(c, fl, i, contrib * ac);
      }
    }
    w_sync();
    double mx = 0.;
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) {
        double v = w.curB[ch * S + s] + seg_sum(w.partA + ch * NM, h.rT_off, s);
        w.curB[ch * S + s] = v;
        mx = fmax(mx, fabs(v));
      }
    const int k = renorm_shift(w_max(mx));
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch)
        t.bO[ch * t.boch + (unsigned)i * S + s] = k ? ldexp(w.curB[ch * S + s], -k) : w.curB[ch * S + s];
    if (lane == 0) t.fO[i] = (double)(fref + k);
    w_sync();
  }
}

// ---- outside, phase EM: E(i,j,s1) <- parent P(i-1,j+1,s) (pair emission at i-1 and j);
//                         M(i,j,s)  <- E(i,j,s) | parent M(i-1,j,sp) emitting x[i-1]
template <int NCH, int MODE = 0>
RDEV void lin_out_EM(const LinCtx& c, const CTabs& t, int i, int d, bool gE, bool gM, WarpLin& w, EhAcc<NCH>& eh) {
  const LinHMM& h = LC.h;
  const LinParams& p = LC.p;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, j = i + d, lane = lane_id(), NM = w.n_max;
  const bool ne = LC.en.no_ene != 0;
  const unsigned il = cidx(q, i, d), ir = cidx(q, j, d);
  double* cE = w.curB;  // [NCH][S] E of this cell
  if (gE) {
    const int xl = q.x[i - 1], xr = q.x[j];
    const double wsl = c.wsf[i - 1], wsr = c.wsf[j];
    const unsigned pb = cidx(q, i - 1, d + 2);
    for (int pz = lane; pz < h.n_pair; pz += WARP_N) {
      int a = ld_ro(h.pT_ord + pz);
      int s = ld_ro(h.p_tgt + a), s1 = ld_ro(h.p_src + a), fl = ld_ro(h.p_flag + a);
      double wt = ld_ro(p.p_w + a * 25 + xl * 5 + xr);
      if (fl & 1) wt *= wsl;
      if (fl & 2) wt *= wsr;
      if ((i - 1 == c.ys && !(fl & LIN_F_START)) || (j == c.ys && !(fl & LIN_F_RSTART))) wt = 0.;
      double ac = t.aE[il + s1];
      // the reference adds to an outside value only when the transition's posterior is non-zero
      // (motif_trainer.hpp:372: `if (zeroL == z) return;` before the updates).  That only shows where its inside and
      // outside passes enumerate different interior loops (LinCtx::Csum): a state of E with no inside derivation must
      // not hand outside weight down to the extra loops.
      if (ac == 0.) wt = 0.;
      for (int ch = 0; ch < NCH; ++ch) {
        double contrib = t.bP[ch * t.bch + pb + s] * wt;
        w.partA[ch * NM + pz] = contrib;
        double post = contrib * ac;
        if (MODE != 2 && !p.no_prf && post != 0.) red_add(w.pcnt + ch * w.pstride + a * 25 + xl * 5 + xr, post);
        if (MODE != 0 && ch == 0 && post != 0.) {
          hook_emit<MODE>(c, fl, i - 1, post);
          hook_emit_right<MODE>(c, fl >> 4, j, post);
        }
      }
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) {
        double v = seg_sum(w.partA + ch * NM, h.pT_off, s);
        cE[ch * S + s] = v;
        t.bEl[ch * t.bch + il + s] = v;
#if !LIN_SCATTER_ILOOP
        t.bEr[ch * t.bch + ir + s] = v;   // only the right-flank gather reads E by its right end
#endif
      }
    w_sync();
  }
  if (gM) {
    bool cM = gE;
    double tM = 0., m0 = 1., m1 = 1.;
    if (gE && !ne) {
      tM = nl_e_ext(&c.q, j, i - 1, 0) + (LC.en.mlclosing + LC.en.mlintern);
      cM = tM > NINF;
      if (cM) { F2 ff = boltz2(tM); m0 = ff.f0; m1 = ff.f1; }
    }
    const bool okM = ok_M(q, i - 1, d + 1);
    if (okM) {
      const int xl = q.x[i - 1];
      const double wsl = c.wsf[i - 1];
      const unsigned pb = cidx(q, i - 1, d + 1);
      for (int pz = lane; pz < h.n_left; pz += WARP_N) {
        int a = ld_ro(h.lT_ord + pz);
        int sp = ld_ro(h.l_tgt + a), s1 = ld_ro(h.l_src + a);
        int fl = ld_ro(h.l_flag + a);
        double wt = ld_ro(p.l_w + a * 5 + xl);
        if (fl & 1) wt *= wsl;
        if (i - 1 == c.ys && !(fl & LIN_F_START)) wt = 0.;
        double ac = t.aM[il + s1];
        for (int ch = 0; ch < NCH; ++ch) {
          double contrib = t.bM[ch * t.bch + pb + sp] * wt;
          w.partA[ch * NM + pz] = contrib;
          if (MODE != 2 && !p.no_prf) w.cntL[(ch * h.n_left + a) * 5 + xl] += contrib * ac;
          if (MODE != 0 && ch == 0 && contrib != 0.) hook_emit<MODE>(c, fl, i - 1, contrib * ac);
        }
      }
      w_sync();
    }
    for (int s = lane; s < S; s += WARP_N) {
      int sl = ld_ro(h.slot + s);
      double am = cM ? t.aM[il + s] : 0.;
      for (int ch = 0; ch < NCH; ++ch) {
        double x = okM ? seg_sum(w.partA + ch * NM, h.lT_off, s) : 0.;
        if (cM) {
          double y = cE[ch * S + s] * (sl ? m1 : m0);
          x += y;
          eh.add(ch, sl, tM * y * am);
        }
        t.bM[ch * t.bch + il + s] = x;
      }
    }
    w_sync();
  }
}

// ---- outside, phase B (cells with gB):
//   1(i,j,sl): left child of B(i,j',s) with right sibling 2(j,j',sr);   B = 1 + M
//   2(i,j,sr): right child of B(i',j,s) with left sibling 1(i',i,sl); + 1(i,j,sr); + parent 2(i,j+1,sp) emitting x[j]
template <int NCH, int MODE = 0> RDEV void lin_out_B(const LinCtx& c, const CTabs& t, int i, int d, bool gM, WarpLin& w) {
  const LinHMM& h = LC.h;
  const LinParams& p = LC.p;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, j = i + d, lane = lane_id(), L = q.L, W = q.W, NM = w.n_max;
  const unsigned il = cidx(q, i, d), ir = cidx(q, j, d);
  double* c1 = w.curB;  // [NCH][S] 1 of this cell
  {
    int nk = 0;
    const unsigned* ri = q.lf + i * q.mw;
    const unsigned* rk = q.lf + j * q.mw;
    int dmax = W < L - i ? W : L - i;
    for (int d20 = d; d20 <= dmax; d20 += WARP_N) {
      int d2 = d20 + lane;
      bool ok = d2 <= dmax && row_bit(ri, d2) && row_bit(rk, d2 - d);
      unsigned bal = w_ballot(ok);
      if (ok) w.kbuf[nk + w_popc(bal & lanemask_lt())] = d2;
      nk += w_popc(bal);
    }
    w_sync();
    // sibling 2(j, j') by right end: row i+d2, span d2-d -> offset ((i+d2)*W1 + d2-d)*S = base + d2*(W1+1)*S
    const int sib0 = (i * q.W1 - d) * S, sstep = (q.W1 + 1) * S;
    const unsigned rb = cidx(q, i, 0);
    for (int pz = lane; pz < h.n_split; pz += WARP_N) {
      int a = ld_ro(h.spL_ord + pz);
      int s = ld_ro(h.sp_tgt + a), sr = ld_ro(h.sp_r + a);
      double v[NCH], v2[NCH];
      for (int ch = 0; ch < NCH; ++ch) { v[ch] = 0.; v2[ch] = 0.; }
      const double* ps = t.a2 + sib0 + sr;
      const double* pb = t.bBl + rb + s;
      int tt = 0;
#if LIN_SPLIT_UNROLL >= 4
      if (NCH == 1) {
        double v3 = 0., v4 = 0.;
        for (; tt + 3 < nk; tt += 4) {   // four split points in flight
          int da = w.kbuf[tt], db = w.kbuf[tt + 1], dc = w.kbuf[tt + 2], de = w.kbuf[tt + 3];
          double sa = ps[da * sstep], sb = ps[db * sstep], sc = ps[dc * sstep], se = ps[de * sstep];
          double ba = pb[(unsigned)da * S], bb = pb[(unsigned)db * S], bc = pb[(unsigned)dc * S], be = pb[(unsigned)de * S];
          v[0] += ba * sa; v2[0] += bb * sb; v3 += bc * sc; v4 += be * se;
        }
        v[0] += v3; v2[0] += v4;
      }
#endif
      for (; tt + 1 < nk; tt += 2) {   // two split points in flight
        int da = w.kbuf[tt], db = w.kbuf[tt + 1];
        double sa = ps[da * sstep], sb = ps[db * sstep];
        for (int ch = 0; ch < NCH; ++ch) {
          v[ch] += pb[ch * t.bch + (unsigned)da * S] * sa;
          v2[ch] += pb[ch * t.bch + (unsigned)db * S] * sb;
        }
      }
      if (tt < nk) {
        int da = w.kbuf[tt];
        double sa = ps[da * sstep];
        for (int ch = 0; ch < NCH; ++ch) v[ch] += pb[ch * t.bch + (unsigned)da * S] * sa;
      }
      for (int ch = 0; ch < NCH; ++ch) w.partA[ch * NM + pz] = v[ch] + v2[ch];
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) {
        double x = seg_sum(w.partA + ch * NM, h.spL_off, s);
        c1[ch * S + s] = x;
        double bb = x + (gM ? t.bM[ch * t.bch + il + s] : 0.);
        t.bBl[ch * t.bch + il + s] = bb;
        t.bBr[ch * t.bch + ir + s] = bb;
      }
    w_sync();
  }
  {
    int nk = 0;
    const unsigned* rj = c.lfr + j * q.mw;
    int dmax = W < j ? W : j;
    for (int d20 = d; d20 <= dmax; d20 += WARP_N) {
      int d2 = d20 + lane;
      bool ok = d2 <= dmax && row_bit(rj, d2) && row_bit(q.lf + (j - d2) * q.mw, d2 - d);
      unsigned bal = w_ballot(ok);
      if (ok) w.kbuf[nk + w_popc(bal & lanemask_lt())] = d2;
      nk += w_popc(bal);
    }
    w_sync();
    // sibling 1(i', i) by left end: row j-d2, span d2-d -> offset ((j-d2)*W1 + d2-d)*S = base - d2*(W1-1)*S
    const int sib0 = (j * q.W1 - d) * S, sstep = (q.W1 - 1) * S;
    const unsigned rb = cidx(q, j, 0);
    for (int pz = lane; pz < h.n_split; pz += WARP_N) {
      int a = ld_ro(h.spR_ord + pz);
      int s = ld_ro(h.sp_tgt + a), sl = ld_ro(h.sp_l + a);
      double v[NCH], v2[NCH];
      for (int ch = 0; ch < NCH; ++ch) { v[ch] = 0.; v2[ch] = 0.; }
      const double* ps = t.a1 + sib0 + sl;
      const double* pb = t.bBr + rb + s;
      int tt = 0;
#if LIN_SPLIT_UNROLL >= 4
      if (NCH == 1) {
        double v3 = 0., v4 = 0.;
        for (; tt + 3 < nk; tt += 4) {
          int da = w.kbuf[tt], db = w.kbuf[tt + 1], dc = w.kbuf[tt + 2], de = w.kbuf[tt + 3];
          double sa = ps[-da * sstep], sb = ps[-db * sstep], sc = ps[-dc * sstep], se = ps[-de * sstep];
          double ba = pb[(unsigned)da * S], bb = pb[(unsigned)db * S], bc = pb[(unsigned)dc * S], be = pb[(unsigned)de * S];
          v[0] += ba * sa; v2[0] += bb * sb; v3 += bc * sc; v4 += be * se;
        }
        v[0] += v3; v2[0] += v4;
      }
#endif
      for (; tt + 1 < nk; tt += 2) {
        int da = w.kbuf[tt], db = w.kbuf[tt + 1];
        double sa = ps[-da * sstep], sb = ps[-db * sstep];
        for (int ch = 0; ch < NCH; ++ch) {
          v[ch] += pb[ch * t.bch + (unsigned)da * S] * sa;
          v2[ch] += pb[ch * t.bch + (unsigned)db * S] * sb;
        }
      }
      if (tt < nk) {
        int da = w.kbuf[tt];
        double sa = ps[-da * sstep];
        for (int ch = 0; ch < NCH; ++ch) v[ch] += pb[ch * t.bch + (unsigned)da * S] * sa;
      }
      for (int ch = 0; ch < NCH; ++ch) w.partA[ch * NM + pz] = v[ch] + v2[ch];
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) c1[ch * S + s] += seg_sum(w.partA + ch * NM, h.spR_off, s);
    w_sync();
    const bool okp = ok_B(q, i, d + 1);
    if (okp) {
      const int xr = q.x[j];
      const double wsr = c.wsf[j];
      const unsigned pb = cidx(q, i, d + 1);
      for (int pz = lane; pz < h.n_right; pz += WARP_N) {
        int a = ld_ro(h.rT_ord + pz);
        int sp = ld_ro(h.r_tgt + a), s1 = ld_ro(h.r_src + a), fl = ld_ro(h.r_flag + a);
        double wt = ld_ro(p.r_w + a * 5 + xr);
        if (fl & 1) wt *= wsr;
        if (j == c.ys && !(fl & LIN_F_START)) wt = 0.;
        double ac = t.a2[ir + s1];
        for (int ch = 0; ch < NCH; ++ch) {
          double contrib = t.b2[ch * t.bch + pb + sp] * wt;
          w.partA[ch * NM + pz] = contrib;
          if (MODE != 2 && !p.no_prf) w.cntR[(ch * h.n_right + a) * 5 + xr] += contrib * ac;
          if (MODE != 0 && ch == 0 && contrib != 0.) hook_emit_right<MODE>(c, fl, j, contrib * ac);
        }
      }
      w_sync();
    }
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) {
        double x = c1[ch * S + s];
        if (okp) x += seg_sum(w.partA + ch * NM, h.rT_off, s);
        t.b2[ch * t.bch + il + s] = x;
      }
    w_sync();
  }
}

// ---- outside, phase P (cells with an allowed pair)
// PART 0: everything.  The wavefront launches PART 1 (2 <- P, stacking parent, exterior parent; stores bP) and PART 2
// (enclosing interior loops; adds to bP) as separate kernels for the same instruction-cache reason as lin_out_L.
template <int NCH, int MODE = 0, int PART = 0>
RDEV void lin_out_P(const LinCtx& c, const CTabs& t, int i, int d, bool gB, WarpLin& w, EhAcc<NCH>& eh) {
  const LinHMM& h = LC.h;
  const LinParams& p = LC.p;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, j = i + d, lane = lane_id(), NM = w.n_max;
  const bool ne = LC.en.no_ene != 0;
  const unsigned il = cidx(q, i, d), ir = cidx(q, j, d);
  double* cP = w.curB;  // [NCH][S]
  if (PART == 2) {
    if (h.n_quad <= 0) return;
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) cP[ch * S + s] = 0.;
    w_sync();
  }
  // 2(i,j,s) <- P(i,j,s)
  if (PART <= 1) {
    bool c2P = gB;
    double tsc = 0., f0 = 1., f1 = 1.;
    if (gB && !ne) {
      tsc = nl_e_ext(&c.q, i, j - 1, 0) + LC.en.mlintern;
      c2P = tsc > NINF;
      if (c2P) { F2 ff = boltz2(tsc); f0 = ff.f0; f1 = ff.f1; }
    }
    for (int s = lane; s < S; s += WARP_N) {
      int sl = ld_ro(h.slot + s);
      double ap = c2P ? t.aP[ir + s] : 0.;
      for (int ch = 0; ch < NCH; ++ch) {
        double y = 0.;
        if (c2P) {
          y = t.b2[ch * t.bch + il + s] * (sl ? f1 : f0);
          eh.add(ch, sl, tsc * y * ap);
        }
        cP[ch * S + s] = y;
      }
    }
    w_sync();
  }
  // parent P(i-1,j+1,s) stacking on this pair
  if (PART <= 1 && ok_P(q, i - 1, d + 2)) {
    bool cPP = true;
    double tsc = 0., f0 = 1., f1 = 1.;
    if (!ne) {
      tsc = nl_e_loop(&c.q, i - 1, j, i, j - 1);
      cPP = tsc > NINF;
      if (cPP) { F2 ff = boltz2(tsc); f0 = ff.f0; f1 = ff.f1; }
    }
    if (cPP) {
      const int xl = q.x[i - 1], xr = q.x[j];
      const double wsl = c.wsf[i - 1], wsr = c.wsf[j];
      const unsigned pb = cidx(q, i - 1, d + 2);
      for (int pz = lane; pz < h.n_pair; pz += WARP_N) {
        int a = ld_ro(h.pT_ord + pz);
        int s = ld_ro(h.p_tgt + a), s1 = ld_ro(h.p_src + a), fl = ld_ro(h.p_flag + a);
        int sl = ld_ro(h.slot + s);
        double wt = ld_ro(p.p_w + a * 25 + xl * 5 + xr) * (sl ? f1 : f0);
        if (fl & 1) wt *= wsl;
        if (fl & 2) wt *= wsr;
        if ((i - 1 == c.ys && !(fl & LIN_F_START)) || (j == c.ys && !(fl & LIN_F_RSTART))) wt = 0.;
        double ac = t.aP[ir + s1];
        for (int ch = 0; ch < NCH; ++ch) {
          double contrib = t.bP[ch * t.bch + pb + s] * wt;
          w.partA[ch * NM + pz] = contrib;
          double post = contrib * ac;
          eh.add(ch, sl, tsc * post);
          if (MODE != 2 && !p.no_prf && post != 0.) red_add(w.pcnt + ch * w.pstride + a * 25 + xl * 5 + xr, post);
          if (MODE != 0 && ch == 0 && post != 0.) {
            hook_emit<MODE>(c, fl, i - 1, post);
            hook_emit_right<MODE>(c, fl >> 4, j, post);
          }
        }
      }
      w_sync();
      for (int s = lane; s < S; s += WARP_N)
        for (int ch = 0; ch < NCH; ++ch) cP[ch * S + s] += seg_sum(w.partA + ch * NM, h.pT_off, s);
      w_sync();
    }
  }
  // exterior parent O(j,s) <- O(i,sl) P(i,j,sr)
  if (PART <= 1) {
    bool cX = true;
    double tsc = 0., f0 = 1., f1 = 1.;
    if (!ne) {
      tsc = nl_e_ext(&c.q, i, j - 1, 1);
      cX = tsc > NINF;
      if (cX) { F2 ff = boltz2(tsc); f0 = ff.f0; f1 = ff.f1; }
    }
    if (cX) {
      // a^O(i) b^O(j): mantissas times 2^(eO(i) + fO(j)), folded into the two Boltzmann factors
      const double xsc = ldexp(1., (int)(t.eO[i] + t.fO[j]));
      f0 *= xsc; f1 *= xsc;
      for (int pz = lane; pz < h.n_split; pz += WARP_N) {
        int a = ld_ro(h.spR_ord + pz);
        int s = ld_ro(h.sp_tgt + a), sl_ = ld_ro(h.sp_l + a), sr = ld_ro(h.sp_r + a);
        int sl = ld_ro(h.slot + s);
        double term = t.aO[(unsigned)i * S + sl_] * (sl ? f1 : f0);
        double ac = t.aP[ir + sr];
        for (int ch = 0; ch < NCH; ++ch) {
          double contrib = t.bO[ch * t.boch + (unsigned)j * S + s] * term;
          w.partA[ch * NM + pz] = contrib;
          eh.add(ch, sl, tsc * contrib * ac);
        }
      }
      w_sync();
      for (int s = lane; s < S; s += WARP_N)
        for (int ch = 0; ch < NCH; ++ch) cP[ch * S + s] += seg_sum(w.partA + ch * NM, h.spR_off, s);
      w_sync();
    }
  }
  // enclosing interior loops E(i',j',s) <- P(i,j,s1) L(i',i,s2) L(j,j',s3)
  if (PART != 1 && h.n_quad > 0) {
    for (int a = lane; a < h.n_quad; a += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) { w.partA[ch * NM + a] = 0.; w.partT[ch * NM + a] = 0.; }
    w_sync();
    walk_outer(c, i, d, w, [&](int n) {
      for (int pz = lane; pz < h.n_quad; pz += WARP_N) {
        int a = ld_ro(h.qP_ord + pz);
        int s = ld_ro(h.q_tgt + a), s2 = ld_ro(h.q_s2 + a), s3 = ld_ro(h.q_s3 + a);
        const double* bf = ld_ro(h.slot + s) ? w.bf1 : w.bf0;
        double v[NCH], vt[NCH];
        for (int ch = 0; ch < NCH; ++ch) { v[ch] = w.partA[ch * NM + pz]; vt[ch] = w.partT[ch * NM + pz]; }
        LIN_PP_PRAGMA
        for (int pp = 0; pp < n; ++pp) {
          int i2 = w.bi[pp], j2 = w.bj[pp];
          double term = t.aLl[cidx(q, i2, i - i2) + s2] * t.aLr[cidx(q, j2, j2 - j) + s3] * bf[pp];
          double tsc = w.bt[pp];
          unsigned eb = cidx(q, i2, j2 - i2) + s;
          for (int ch = 0; ch < NCH; ++ch) {
            double x = t.bEl[ch * t.bch + eb] * term;
            v[ch] += x;
            vt[ch] += tsc * x;
          }
        }
        for (int ch = 0; ch < NCH; ++ch) { w.partA[ch * NM + pz] = v[ch]; w.partT[ch * NM + pz] = vt[ch]; }
      }
    });
    for (int pz = lane; pz < h.n_quad; pz += WARP_N) {
      int a = ld_ro(h.qP_ord + pz);
      int sl = ld_ro(h.slot + ld_ro(h.q_tgt + a));
      double ac = t.aP[ir + ld_ro(h.q_s1 + a)];
      for (int ch = 0; ch < NCH; ++ch) eh.add(ch, sl, w.partT[ch * NM + pz] * ac);
    }
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) cP[ch * S + s] += seg_sum(w.partA + ch * NM, h.qP_off, s);
    w_sync();
  }
  if (PART == 2 || (LIN_SCATTER_ILOOP && PART == 1)) {
    // scatter mode: the enclosing loops have already added their share (lin_out_ES at larger spans) to the zeroed entry
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) t.bP[ch * t.bch + il + s] += cP[ch * S + s];
  } else {
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) t.bP[ch * t.bch + il + s] = cP[ch * S + s];
  }
  w_sync();
}

// ---- outside, phase L (every cell): hairpin E(i,j,s) <- L(i,j,s); parent L(i,j+1,sp) emitting x[j]; unpaired flanks
// PART 0: everything in one call.  The wavefront launches it as three kernels -- PART 1 (hairpin + parent L, stores
// bL), PART 2 (left flanks, adds to bL), PART 3 (right flanks, adds to bL) -- because the fused body is ~4 400 SASS
// instructions (70 KB) against a 32 KB instruction cache and a third of its stall cycles were instruction fetch.  The
// order of the floating-point additions is the same either way.
template <int NCH, int MODE = 0, int PART = 0>
RDEV void lin_out_L(const LinCtx& c, const CTabs& t, int i, int d, bool gE, WarpLin& w, EhAcc<NCH>& eh) {
  const LinHMM& h = LC.h;
  const LinParams& p = LC.p;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, j = i + d, lane = lane_id(), L = q.L, W = q.W, NM = w.n_max;
  const bool ne = LC.en.no_ene != 0;
  const unsigned il = cidx(q, i, d);
  double* cL = w.curB;  // [NCH][S]
  if (PART >= 2) {
    if (!(d >= 1 && d <= c.Ceff && h.n_quad > 0)) return;
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) cL[ch * S + s] = 0.;
    w_sync();
  }
  if (PART <= 1) {
    bool cH = gE;
    double tH = 0., h0 = 1., h1 = 1.;
    if (gE && !ne) {
      tH = nl_e_hairpin(&c.q, i - 1, j);
      cH = tH > NINF;
      if (cH) { F2 ff = boltz2(tH); h0 = ff.f0; h1 = ff.f1; }
    }
    for (int s = lane; s < S; s += WARP_N) {
      int sl = ld_ro(h.slot + s);
      bool use = cH && ld_ro(h.is_loop + s);
      double al = use ? t.aLl[il + s] : 0.;
      for (int ch = 0; ch < NCH; ++ch) {
        double y = 0.;
        if (use) {
          y = t.bEl[ch * t.bch + il + s] * (sl ? h1 : h0);
          eh.add(ch, sl, tH * y * al);
        }
        cL[ch * S + s] = y;
      }
    }
    w_sync();
  }
  if (PART <= 1 && d + 1 <= W && j + 1 <= L) {
    const int xr = q.x[j];
    const double wsr = c.wsf[j];
    const unsigned pb = cidx(q, i, d + 1);
    for (int pz = lane; pz < h.n_right; pz += WARP_N) {
      int a = ld_ro(h.rT_ord + pz);
      int fl = ld_ro(h.r_flag + a);
      int sp = ld_ro(h.r_tgt + a), s1 = ld_ro(h.r_src + a);
      double wt = 0.;
      if (fl & 2) {
        wt = ld_ro(p.r_w + a * 5 + xr);
        if (fl & 1) wt *= wsr;
        if (j == c.ys && !(fl & LIN_F_START)) wt = 0.;
      }
      double ac = t.aLl[il + s1];
      for (int ch = 0; ch < NCH; ++ch) {
        double contrib = (fl & 2) ? t.bL[ch * t.bch + pb + sp] * wt : 0.;
        w.partA[ch * NM + pz] = contrib;
        if (MODE != 2 && !p.no_prf) w.cntR[(ch * h.n_right + a) * 5 + xr] += contrib * ac;
        if (MODE != 0 && ch == 0 && contrib != 0.) hook_emit_right<MODE>(c, fl, j, contrib * ac);
      }
    }
    w_sync();
    for (int s = lane; s < S; s += WARP_N)
      for (int ch = 0; ch < NCH; ++ch) cL[ch * S + s] += seg_sum(w.partA + ch * NM, h.rT_off, s);
    w_sync();
  }
  bool touched = false;
  if (PART != 1 && d >= 1 && d <= c.Ceff && h.n_quad > 0) {
    if (PART != 3 && i >= 1) {
      const unsigned eb = cidx(q, i, 0);
      bool any = false;
      walk_left_flank(c, i, d, w, [&](int n) {
        if (!any) {   // most cells flank no interior loop at all: set the accumulators up only when one shows up
          for (int a = lane; a < h.n_quad; a += WARP_N)
            for (int ch = 0; ch < NCH; ++ch) w.partA[ch * NM + a] = 0.;
          w_sync();
          any = true;
        }
        for (int pz = lane; pz < h.n_quad; pz += WARP_N) {
          int a = ld_ro(h.qL_ord + pz);
          int s = ld_ro(h.q_tgt + a), s1 = ld_ro(h.q_s1 + a), s3 = ld_ro(h.q_s3 + a);
          const double* bf = ld_ro(h.slot + s) ? w.bf1 : w.bf0;
          double v[NCH];
          for (int ch = 0; ch < NCH; ++ch) v[ch] = w.partA[ch * NM + pz];
          LIN_PP_PRAGMA
        for (int pp = 0; pp < n; ++pp) {
            int l = w.bi[pp], j2 = w.bj[pp];
            double term = t.aP[cidx(q, l, l - j) + s1] * t.aLr[cidx(q, j2, j2 - l) + s3] * bf[pp];
            for (int ch = 0; ch < NCH; ++ch) v[ch] += t.bEl[ch * t.bch + eb + (unsigned)(j2 - i) * S + s] * term;
          }
          for (int ch = 0; ch < NCH; ++ch) w.partA[ch * NM + pz] = v[ch];
        }
      });
      if (any)
        for (int s = lane; s < S; s += WARP_N)
          for (int ch = 0; ch < NCH; ++ch) cL[ch * S + s] += seg_sum(w.partA + ch * NM, h.qL_off, s);
      touched = touched || any;
      w_sync();
    }
    if (PART != 2 && j + 1 <= L) {
      const unsigned eb = cidx(q, j, 0);
      bool any = false;
      walk_right_flank(c, i, d, w, [&](int n) {
        if (!any) {
          for (int a = lane; a < h.n_quad; a += WARP_N)
            for (int ch = 0; ch < NCH; ++ch) w.partA[ch * NM + a] = 0.;
          w_sync();
          any = true;
        }
        for (int pz = lane; pz < h.n_quad; pz += WARP_N) {
          int a = ld_ro(h.qR_ord + pz);
          int s = ld_ro(h.q_tgt + a), s1 = ld_ro(h.q_s1 + a), s2 = ld_ro(h.q_s2 + a);
          const double* bf = ld_ro(h.slot + s) ? w.bf1 : w.bf0;
          double v[NCH];
          for (int ch = 0; ch < NCH; ++ch) v[ch] = w.partA[ch * NM + pz];
          LIN_PP_PRAGMA
        for (int pp = 0; pp < n; ++pp) {
            int k = w.bi[pp], i2 = w.bj[pp];
            double term = t.aP[cidx(q, i, i - k) + s1] * t.aLl[cidx(q, i2, k - i2) + s2] * bf[pp];
            for (int ch = 0; ch < NCH; ++ch) v[ch] += t.bEr[ch * t.bch + eb + (unsigned)(j - i2) * S + s] * term;
          }
          for (int ch = 0; ch < NCH; ++ch) w.partA[ch * NM + pz] = v[ch];
        }
      });
      if (any)
        for (int s = lane; s < S; s += WARP_N)
          for (int ch = 0; ch < NCH; ++ch) cL[ch * S + s] += seg_sum(w.partA + ch * NM, h.qR_off, s);
      touched = touched || any;
      w_sync();
    }
  }
  if (PART >= 2) {
    if (touched)
      for (int s = lane; s < S; s += WARP_N)
        for (int ch = 0; ch < NCH; ++ch) t.bL[ch * t.bch + il + s] += cL[ch * S + s];
  } else if (d >= 1) {
    if (LIN_SCATTER_ILOOP && PART == 1 && d <= c.Cfl && h.n_quad > 0) {
      // a possible flank: the loops it flanks have already added their share (lin_out_ES) to the zeroed entry
      for (int s = lane; s < S; s += WARP_N)
        for (int ch = 0; ch < NCH; ++ch) t.bL[ch * t.bch + il + s] += cL[ch * S + s];
    } else {
      for (int s = lane; s < S; s += WARP_N)
        for (int ch = 0; ch < NCH; ++ch) t.bL[ch * t.bch + il + s] = cL[ch * S + s];
    }
  }
  w_sync();
}

// ---- outside, phase ES (cells enclosed by a pair; scatter mode): E(i,j,s) <- P(k,l,s1) L(i,k,s2) L(l,j,s3).
// With b^E(i,j,.) final, every (inner pair, quad) term t = b^E f a^P a^L a^L is a transition posterior; the outside values
// of its children are t divided by the child's own inside value, i.e. the product of the other factors.
template <int NCH, int MODE = 0>
RDEV void lin_out_ES(const LinCtx& c, const CTabs& t, int i, int d, WarpLin& w, EhAcc<NCH>& eh) {
  const LinHMM& h = LC.h;
  const SeqView q = c.q;  // private copy: the shared-memory original would be re-read after every shared store
  const int S = q.S, j = i + d, lane = lane_id();
  if (h.n_quad <= 0) return;
  const unsigned il = cidx(q, i, d);
  const unsigned r0 = cidx(q, i, 0);
  const double* rL = t.aLl + r0;
  const double* rR = t.aLr + cidx(q, j, 0);
  walk_inner(c, i, d, w, c.Csum, [&](int n) {
    for (int a = lane; a < h.n_quad; a += WARP_N) {
      const int s = ld_ro(h.q_tgt + a), s1 = ld_ro(h.q_s1 + a), s2 = ld_ro(h.q_s2 + a), s3 = ld_ro(h.q_s3 + a);
      const int sl = ld_ro(h.slot + s);
      const double* bf = sl ? w.bf1 : w.bf0;
      double be[NCH];
      bool any = false;
      for (int ch = 0; ch < NCH; ++ch) { be[ch] = ld_ro(t.bEl + ch * t.bch + il + s); any = any || be[ch] != 0.; }
      if (!any) continue;
      // four inner pairs at a time, all loads first: the REDs below may alias anything as far as the compiler knows,
      // so a load issued after one waits for it -- with one pair per iteration every pair paid a full memory latency
      // (the inside tables and b^E are not written by this kernel: read-only loads)
      for (int p0 = 0; p0 < n; p0 += 4) {
        double ap[4], al[4], ar[4], f[4];
        int kk[4], ll[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int pp = p0 + u;
          const bool ok = pp < n;
          kk[u] = ok ? w.bi[pp] : i; ll[u] = ok ? w.bj[pp] : j;
          f[u] = ok ? bf[pp] : 0.;
          ap[u] = ld_ro(t.aP + cidx(q, ll[u], ll[u] - kk[u]) + s1);
          al[u] = ld_ro(rL + (unsigned)(kk[u] - i) * S + s2);
          ar[u] = ld_ro(rR + (unsigned)(j - ll[u]) * S + s3);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (f[u] == 0.) continue;
          const int k = kk[u], l = ll[u];
          const double lr = al[u] * ar[u], tsc = w.bt[p0 + u];
          const unsigned ip = cidx(q, k, l - k) + s1, ifl = r0 + (unsigned)(k - i) * S + s2, ifr = cidx(q, l, j - l) + s3;
          for (int ch = 0; ch < NCH; ++ch) {
            const double x = be[ch] * f[u];
            if (x == 0.) continue;
            const double xp = x * ap[u];
            if (lr != 0.) red_add(t.bP + ch * t.bch + ip, x * lr);
            if (k > i && xp * ar[u] != 0.) red_add(t.bL + ch * t.bch + ifl, xp * ar[u]);
            if (l < j && xp * al[u] != 0.) red_add(t.bL + ch * t.bch + ifr, xp * al[u]);
            eh.add(ch, sl, tsc * (xp * lr));
          }
        }
      }
    }
  });
}


}  // namespace lin
}  // namespace relem
#endif
