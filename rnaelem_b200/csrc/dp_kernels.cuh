// dp_kernels.cuh -- the persistent kernels: one CTA per sequence slot, sequences pulled from a queue.
//   relem_estep_kernel   K0 (energy-only base-pair filter) + coupled inside + 2-channel outside with counts
//                        (RNAelemTrainDP::operator(), motif_trainer.hpp:204-245)
//   relem_scan_kernel    K0 + inside + outside(start/inner posteriors) + start-constrained inside/outside
//                        (end posteriors) + constrained Viterbi + traceback (RNAelemScanDP, motif_scanner.hpp:172-260)
//   relem_bpp_kernel     K0 alone (EnergyModel::fill_bpp_tables, energy_model.hpp:211-266)
#ifndef RELEM_DP_KERNELS_CUH
#define RELEM_DP_KERNELS_CUH
#include "dp_batch.hpp"
#include "dp_warp.cuh"
#include "dp_vit.cuh"

namespace relem {
namespace dp {

#define RELEM_CTA_THREADS 128
#ifndef RELEM_VIT_THREADS
#define RELEM_VIT_THREADS 256   // Viterbi kernel: 8 warps share one sequence (measured against 64: profiles/r2_viterbi.md)
#endif

// per-slot scratch, offsets in doubles from the slot base
struct SlotLayout {
  unsigned long long stride;
  unsigned long long tabA, Q0, Q1, tab0, q0, otab, QO0, QO1, otab0, QO00, emit0, emitT, zeros, G, stack;
  int Lmax, Wmax, mw;
  // dynamic shared memory carve-up (byte offsets)
  int sm_x, sm_sp3, sm_sp4, sm_sp6, sm_bp, sm_lf, sm_bp2, sm_en, sm_pys, sm_pyi, sm_pye, sm_red, sm_ctr, sm_warp, warp_bytes,
      sm_total;
  // Viterbi kernel (its own carve-up, see make_layout): claim slot, VitSeq[R], R per-sequence blocks (bases, special
  // hairpins, two masks), one VitWarp slice per warp
  int sm_vit_claim, sm_vit_ctx, sm_vit_seq, vit_seq_bytes, vit_R, sm_vit_warp, vit_warp_bytes, sm_total_vit, vit_n_max;
};

struct BppOut {
  const long long* moff;  // [nseq+1] byte offsets of the per-sequence masks
  unsigned char* bp_ok; unsigned char* left_ok; double* lnbpp; double* bpp_eff; double* lnZ;
};

struct Smem {
  unsigned char* x; signed char *sp3, *sp4, *sp6;
  unsigned *bp, *lf, *bp2;
  double *en, *pys, *pyi, *pye, *red;
};
RDEV Smem carve(unsigned char* base, const SlotLayout& lay) {
  Smem s;
  s.x = base + lay.sm_x;
  s.sp3 = (signed char*)(base + lay.sm_sp3); s.sp4 = (signed char*)(base + lay.sm_sp4);
  s.sp6 = (signed char*)(base + lay.sm_sp6);
  s.bp = (unsigned*)(base + lay.sm_bp); s.lf = (unsigned*)(base + lay.sm_lf); s.bp2 = (unsigned*)(base + lay.sm_bp2);
  s.en = (double*)(base + lay.sm_en); s.pys = (double*)(base + lay.sm_pys); s.pyi = (double*)(base + lay.sm_pyi);
  s.pye = (double*)(base + lay.sm_pye); s.red = (double*)(base + lay.sm_red);
  return s;
}

RDEV void cta_zero(double* p, unsigned long long n) {
  for (unsigned long long t = CTA_TID; t < n; t += CTA_NTH) p[t] = 0.;
}

// block-wide sum of one double per thread; result valid in every thread.  red: >= CTA_NTH doubles? no: 33.
RDEV double cta_sum(double v, double* red) {
#ifdef RELEM_HOST_EMU
  (void)red;
  return v;
#else
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  int w = CTA_TID >> 5, nw = (CTA_NTH + 31) >> 5;
  CTA_SYNC();
  if ((CTA_TID & 31) == 0) red[w] = v;
  CTA_SYNC();
  double r = 0.;
  for (int k = 0; k < nw; ++k) r += red[k];
  CTA_SYNC();
  return r;
#endif
}

// Sequence set-up shared by all kernels: loads the bases, builds masks, runs the energy-only filter.
// Returns bpp_eff; on return sm.bp / sm.lf hold the masks the coupled passes use.
RDEV double cta_prepare(const ModelView& nullm, const ModelView& m, const BatchView& b, int n, const SlotLayout& lay,
                        double* slot, Smem& sm, SeqView& q, double* lnbpp_out, double* lnZ_out) {
  long long o = b.off[n];
  int L = (int)(b.off[n + 1] - o);
  int W = L < m.en.max_span ? L : m.en.max_span;
  int C = W - 7 < m.en.max_iloop ? W - 7 : m.en.max_iloop;
  q.L = L; q.W = W; q.C = C; q.W1 = W + 1; q.cells = (unsigned)(L + 1) * (unsigned)(W + 1);
  q.mw = lay.mw; q.min_pair = 5; q.min_multi = 10; q.bpr = nullptr; q.lfr = nullptr;
  q.x = sm.x; q.bp = sm.bp; q.lf = sm.lf; q.sp3 = sm.sp3; q.sp4 = sm.sp4; q.sp6 = sm.sp6;
  q.ws = b.ws + o;
  for (int t = CTA_TID; t < L; t += CTA_NTH) sm.x[t] = b.seq[o + t];
  if (CTA_TID == 0) sm.x[L] = 0;
  CTA_SYNC();
  cta_special_hairpins(m.en, sm.x, L, sm.sp3, sm.sp4, sm.sp6);
  cta_canonical_mask(q, sm.bp);
  CTA_SYNC();
  cta_left_mask(q, sm.bp, sm.lf);
  int total = cta_count_bits(sm.bp, (L + 1) * lay.mw, (int*)sm.red);
  int nbp = total;
  if (m.en.filter) {
    // energy-only inside/outside with the one-state null automaton (EnergyModel::calc_BPP, energy_model.hpp:188-193)
    SeqView q0 = q;
    q0.S = 1; q0.emit0 = slot + lay.zeros; q0.emitT = slot + lay.zeros;
    double* tab0 = slot + lay.tab0;
    double* Q0 = slot + lay.q0;
    double* otab0 = slot + lay.otab0;
    double* QO00 = slot + lay.QO00;
    cta_zero(slot + lay.zeros, (unsigned long long)(L > 0 ? L : 1));
    cta_zero(Q0, (unsigned long long)NPLANE * q0.cells);
    cta_zero(QO00, (unsigned long long)(L + 1));
    CTA_SYNC();
    cta_inside<false, NoConstraint>(nullm, q0, tab0, otab0, nullptr, nullptr, NoConstraint());
    double root[3] = {1., 0., 0.};
    if (lnZ_out && CTA_TID == 0) *lnZ_out = otab0[L];
    Counts cn; cn.G = nullptr; cn.ENp = nullptr; cn.Pys = cn.Pyi = cn.Pye = nullptr; cn.n_theta = 0; cn.ML = 0;
    double eh[2];
    cta_outside<1, HOOK_NONE, NoConstraint>(nullm, q0, tab0, otab0, Q0, nullptr, QO00, nullptr, root, NoConstraint(),
                                            cn, eh);
    // keep pairs with ln BPP >= ln min_bpp (energy_model.hpp:257-261)
    for (int t = CTA_TID; t < (L + 1) * lay.mw; t += CTA_NTH) {
      int i = t / lay.mw, w = t % lay.mw;
      unsigned in = sm.bp[t], out = 0u;
      for (int bb = 0; bb < 32; ++bb) {
        if (!((in >> bb) & 1u)) continue;
        int d = w * 32 + bb;
        double qp = ld_cg(Q0 + band_idx(q0, PL_P, i, d, 0));
        double ln = qp > 0. ? log(qp) : NINF;
        if (lnbpp_out) lnbpp_out[i * q.W1 + d] = ln;
        if (m.en.min_lnbpp <= ln) out |= 1u << bb;
      }
      sm.bp2[t] = out;
    }
    CTA_SYNC();
    for (int t = CTA_TID; t < (L + 1) * lay.mw; t += CTA_NTH) sm.bp[t] = sm.bp2[t];
    CTA_SYNC();
    cta_left_mask(q, sm.bp, sm.lf);
    nbp = cta_count_bits(sm.bp, (L + 1) * lay.mw, (int*)sm.red);
  }
  CTA_SYNC();
  return (double)nbp / (double)total;
}

// G (node, position) posteriors -> theta-shaped counts, added to the shared EN accumulators
RDEV void cta_fold_G(const ModelView& m, const SeqView& q, const double* G, double* en, int nch) {
  const DevHMM& h = m.h;
  int ML = h.M * q.L;
  for (int t = CTA_TID; t < ML; t += CTA_NTH) {
    int hn = t / q.L, p = t % q.L;
    int tid = ld_ro(h.theta_id + hn);
    int bse = q.x[p];
    if (tid < 0 || bse == 0) continue;
    int idx = ld_ro(h.theta_off + tid) + bse - 1;
    for (int c = 0; c < nch; ++c) {
      double g = ld_cg(G + c * ML + t);
      if (g != 0.) red_add(en + c * m.p.n_theta + idx, g);
    }
  }
}

#ifdef RELEM_HOST_EMU
#define PROF_MARK(slot) ((void)0)
#else
// phase clock: thread 0 adds the cycles since the previous mark to prof[slot]
#define PROF_MARK(slot)                                                        \
  do {                                                                         \
    if (CTA_TID == 0 && out.prof) {                                            \
      long long now__ = clock64();                                             \
      atomicAdd(out.prof + (slot), (unsigned long long)(now__ - prof_t0));     \
      prof_t0 = now__;                                                         \
    }                                                                          \
  } while (0)
#endif

RDEV int claim(int* queue, int* sh) {
  if (CTA_TID == 0) {
#ifdef RELEM_HOST_EMU
    *sh = (*queue)++;
#else
    *sh = atomicAdd(queue, 1);
#endif
  }
  CTA_SYNC();
  int r = *sh;
  CTA_SYNC();
  return r;
}

#ifdef RELEM_HOST_EMU
#define RELEM_KERNEL inline void
#define RELEM_SMEM_DECL unsigned char* smem_raw
#define RELEM_SMEM_ARG , unsigned char* smem_raw
#define RELEM_BLOCK_IDX 0
#else
#define RELEM_KERNEL __global__ void __launch_bounds__(RELEM_CTA_THREADS)
#define RELEM_SMEM_ARG
#define RELEM_BLOCK_IDX ((int)blockIdx.x)
#endif

// ------------------------------------------------------------------------------------------------ E-step
RELEM_KERNEL relem_estep_kernel(ModelView nullm, ModelView m, BatchView b, SlotLayout lay, double* scratch,
                                int* queue, EstepOut out RELEM_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  Smem sm = carve(smem_raw, lay);
  double* slot = scratch + (unsigned long long)RELEM_BLOCK_IDX * lay.stride;
  const int S = m.h.S, NT = m.p.n_theta;
  WarpSm wsm = warp_carve(smem_raw + lay.sm_warp + warp_id() * lay.warp_bytes, S, lay.Wmax);
  int* ctr = (int*)(smem_raw + lay.sm_ctr);
  if (CTA_TID == 0) { ctr[0] = 0; ctr[1] = 0; }
  CTA_SYNC();
  for (;;) {
    int qi = claim(queue, (int*)(sm.red + 40));
    if (qi >= b.nseq) break;
    int n = b.order[qi];
    SeqView q;
    q.S = S;
#ifndef RELEM_HOST_EMU
    long long prof_t0 = clock64();
#endif
    double eff = cta_prepare(nullm, m, b, n, lay, slot, sm, q, nullptr, nullptr);
    PROF_MARK(0);
    double* emit0 = slot + lay.emit0; double* emitT = slot + lay.emitT;
    q.emit0 = emit0; q.emitT = emitT;
    cta_emit_tables(m, q, emit0, emitT);
    CTA_SYNC();
    double* tab = slot + lay.tabA; double* otab = slot + lay.otab;
    double* Q0 = slot + lay.Q0; double* Q1 = slot + lay.Q1;
    unsigned long long nt = (unsigned long long)NPLANE * q.cells * S;
    cta_inside_warp(m, q, tab, otab, NoConstraint(), wsm, ctr, Q0, Q1, nt);
    PROF_MARK(1);
    int L = q.L;
    double Ztt = part_func(m.h, otab, L, S, true, true);
    double Ztf = part_func(m.h, otab, L, S, true, false);
    double Zft = part_func(m.h, otab, L, S, false, true);
    int kind = b.kind[n];
    bool fin_tt = Ztt > NINF && Ztt < -NINF, fin_tf = Ztf > NINF && Ztf < -NINF;
    bool skip = kind == 2 ? !fin_tt : !(fin_tt && fin_tf);
    if (CTA_TID == 0) {
      out.Z[n * 3 + 0] = Ztt; out.Z[n * 3 + 1] = Ztf; out.Z[n * 3 + 2] = Zft;
      out.bpp_eff[n] = eff; out.skipped[n] = skip ? 1 : 0;
    }
    if (skip) {
      for (int t = CTA_TID; t < NT; t += CTA_NTH) { out.ENo[(long long)n * NT + t] = 0.; out.ENx[(long long)n * NT + t] = 0.; }
      for (int t = CTA_TID; t < 4; t += CTA_NTH) out.EH[n * 4 + t] = 0.;
      CTA_SYNC();
      continue;
    }
    double* QO0 = slot + lay.QO0; double* QO1 = slot + lay.QO1;
    double* G = slot + lay.G;
    cta_zero(QO0, (unsigned long long)(L + 1) * S); cta_zero(QO1, (unsigned long long)(L + 1) * S);
    cta_zero(G, 2ull * m.h.M * L);
    cta_zero(sm.en, 2ull * NT);
    CTA_SYNC();
    PROF_MARK(2);
    // channel 0: both boundary states open (Zo); channel 1: the restricted condition (Zx)
    double root[6];
    const DevHMM& h = m.h;
    double r00 = h.s00 >= 0 ? otab[L * S + h.s00] : NINF;
    double rM2 = h.s0M2 >= 0 ? otab[L * S + h.s0M2] : NINF;
    double rM1 = h.s0M1 >= 0 ? otab[L * S + h.s0M1] : NINF;
    root[0] = exp(r00 - Ztt); root[1] = exp(rM2 - Ztt); root[2] = exp(rM1 - Ztt);
    if (kind == 1 || kind >= 3) { root[3] = 0.; root[4] = exp(rM2 - Ztf); root[5] = exp(rM1 - Ztf); }
    else { root[3] = (Zft > NINF) ? exp(r00 - Zft) : 0.; root[4] = 0.; root[5] = 0.; }
    Counts cn; cn.G = G; cn.ENp = sm.en; cn.Pys = cn.Pyi = cn.Pye = nullptr; cn.n_theta = NT; cn.ML = m.h.M * L;
    double eh[4];
    cta_outside_warp<2, HOOK_TRAIN, false>(m, q, tab, otab, Q0, Q1, QO0, QO1, root, NoConstraint(), cn, eh, wsm, ctr);
    CTA_SYNC();
    PROF_MARK(3);
    if (!m.p.no_prf) cta_fold_G(m, q, G, sm.en, 2);
    double e0 = cta_sum(eh[0], sm.red), e1 = cta_sum(eh[1], sm.red), e2 = cta_sum(eh[2], sm.red),
           e3 = cta_sum(eh[3], sm.red);
    CTA_SYNC();
    for (int t = CTA_TID; t < NT; t += CTA_NTH) {
      out.ENo[(long long)n * NT + t] = sm.en[t];
      out.ENx[(long long)n * NT + t] = sm.en[NT + t];
    }
    if (CTA_TID == 0) { out.EH[n * 4 + 0] = e0; out.EH[n * 4 + 1] = e1; out.EH[n * 4 + 2] = e2; out.EH[n * 4 + 3] = e3; }
    CTA_SYNC();
    PROF_MARK(4);
  }
}

// batch reduction (the per-thread accumulate + mutex block of motif_trainer.hpp:248-271, made deterministic):
// res = [fn, sum_eff, n_skipped, EHo0, EHo1, EHx0, EHx1, ENo[NT], ENx[NT]]
// step 1: a gated sequence (the negative of a skipped positive) is dropped with its gate
RELEM_KERNEL relem_gate_kernel(int nseq, const int* gate, EstepOut out RELEM_SMEM_ARG) {
#ifdef RELEM_HOST_EMU
  (void)smem_raw;
  for (int n = 0; n < nseq; ++n) {
#else
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nseq; n += gridDim.x * blockDim.x) {
#endif
    int g = gate ? gate[n] : -1;
    if (g >= 0 && out.skipped[g] == 1) out.skipped[n] = 2;
  }
}
// step 2: one CTA per output; thread p sums sequences p, p+T, ... and a fixed-shape tree adds the partial sums, so
// the result does not depend on scheduling
RELEM_KERNEL relem_reduce_kernel(int nseq, int NT, const unsigned char* kind, EstepOut out, double* res,
                                 int emu_block RELEM_SMEM_ARG) {
#ifdef RELEM_HOST_EMU
  (void)smem_raw;
  const int t = emu_block;
  double red1[1];
  double* red = red1;
#else
  (void)emu_block;
  const int t = (int)blockIdx.x;
  __shared__ double red[RELEM_CTA_THREADS];
#endif
  double acc = 0.;
  for (int n = CTA_TID; n < nseq; n += CTA_NTH) {
    int sk = out.skipped[n];
    if (t == 2) { if (sk == 1) acc += 1.; continue; }
    if (sk) continue;
    int kd = kind[n];
    // --lik-ratio kinds 3 / 4 (motif_trainer.hpp:156-202): the pair Z(1,1) / Z(1,0) with the roles of Zo and Zx
    // swapped, i.e. the kind-1 terms with the opposite sign
    const bool with_pair = kd == 1 || kd >= 3;
    const double sg = kd >= 3 ? -1. : 1.;
    if (t == 0) acc += sg * (out.Z[n * 3 + 0] - (with_pair ? out.Z[n * 3 + 1] : out.Z[n * 3 + 2]));
    else if (t == 1) { if (kd != 2 && kd != 4) acc += out.bpp_eff[n]; }
    else if (t < 7) acc += sg * out.EH[n * 4 + (t - 3)];
    else if (t < 7 + NT) acc += sg * out.ENo[(long long)n * NT + (t - 7)];
    else acc += sg * out.ENx[(long long)n * NT + (t - 7 - NT)];
  }
  red[CTA_TID] = acc;
  CTA_SYNC();
  for (int o = CTA_NTH / 2; o > 0; o >>= 1) {
    if (CTA_TID < o) red[CTA_TID] += red[CTA_TID + o];
    CTA_SYNC();
  }
  if (CTA_TID == 0) res[t] = red[0];
}

// -------------------------------------------------------------------------------------------------- bpp
RELEM_KERNEL relem_bpp_kernel(ModelView nullm, ModelView m, BatchView b, SlotLayout lay, double* scratch, int* queue,
                              BppOut out RELEM_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  Smem sm = carve(smem_raw, lay);
  double* slot = scratch + (unsigned long long)RELEM_BLOCK_IDX * lay.stride;
  for (;;) {
    int qi = claim(queue, (int*)(sm.red + 40));
    if (qi >= b.nseq) break;
    int n = b.order[qi];
    SeqView q;
    q.S = m.h.S;
    long long mo = out.moff[n];
    int L = (int)(b.off[n + 1] - b.off[n]);
    int W = L < m.en.max_span ? L : m.en.max_span;
    double* ln = out.lnbpp ? out.lnbpp + mo : nullptr;
    if (ln) {
      for (int t = CTA_TID; t < (L + 1) * (W + 1); t += CTA_NTH) ln[t] = NINF;
      CTA_SYNC();
    }
    double eff = cta_prepare(nullm, m, b, n, lay, slot, sm, q, ln, out.lnZ ? out.lnZ + n : nullptr);
    if (CTA_TID == 0 && out.bpp_eff) out.bpp_eff[n] = eff;
    for (int t = CTA_TID; t < (L + 1) * (W + 1); t += CTA_NTH) {
      int i = t / (W + 1), d = t % (W + 1);
      if (out.bp_ok) out.bp_ok[mo + t] = mask_bit(sm.bp, lay.mw, i, d) ? 1 : 0;
      if (out.left_ok) out.left_ok[mo + t] = mask_bit(sm.lf, lay.mw, i, d) ? 1 : 0;
    }
    CTA_SYNC();
  }
}

// ------------------------------------------------------------------------------------------------- scan
// index of the last maximum (max_index, util.hpp:231-241), NaN never wins
RDEV int last_max_index(const double* v, int n) {
  int s = 0;
  double mx = -1.7976931348623157e308;
  for (int i = 0; i < n; ++i)
    if (mx <= v[i]) { s = i; mx = v[i]; }
  return s;
}


// RNAelemScanDP::trace_back (motif_scanner.hpp:262-362), one thread
RDEV void trace_back(const ModelView& m, const SeqView& q, const unsigned long long* trace,
                     const unsigned long long* otrace, const int* n2s, int* stack, int s0, int* psihat, char* rss) {
  const DevHMM& h = m.h;
  const int S = q.S, M = h.M;
  int sp = 0;
#define PUSH(I, J, E, SS) { stack[sp * 4] = (I); stack[sp * 4 + 1] = (J); stack[sp * 4 + 2] = (E); stack[sp * 4 + 3] = (SS); ++sp; }
  PUSH(0, q.L, 7, s0)
  while (sp > 0) {
    --sp;
    int ti = stack[sp * 4], tj = stack[sp * 4 + 1], te = stack[sp * 4 + 2], ts = stack[sp * 4 + 3];
    unsigned long long tr = (te == 7) ? otrace[tj * S + ts] : trace[band_idx(q, te, ti, tj - ti, ts)];
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

RELEM_KERNEL relem_scan_kernel(ModelView nullm, ModelView m, BatchView b, SlotLayout lay, double* scratch, int* queue,
                               const int* n2s, ScanOut out RELEM_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  Smem sm = carve(smem_raw, lay);
  double* slot = scratch + (unsigned long long)RELEM_BLOCK_IDX * lay.stride;
  const int S = m.h.S, NT = m.p.n_theta;
  const DevHMM& h = m.h;
  WarpSm wsm = warp_carve(smem_raw + lay.sm_warp + warp_id() * lay.warp_bytes, S, lay.Wmax);
  int* ctr = (int*)(smem_raw + lay.sm_ctr);
  if (CTA_TID == 0) { ctr[0] = 0; ctr[1] = 0; }
  CTA_SYNC();
  for (;;) {
    int qi = claim(queue, (int*)(sm.red + 40));
    if (qi >= b.nseq) break;
    int n = b.order[qi];
    SeqView q;
    q.S = S;
    cta_prepare(nullm, m, b, n, lay, slot, sm, q, nullptr, nullptr);
    double* emit0 = slot + lay.emit0; double* emitT = slot + lay.emitT;
    q.emit0 = emit0; q.emitT = emitT;
    cta_emit_tables(m, q, emit0, emitT);
    CTA_SYNC();
    const int L = q.L;
    long long o = b.off[n];
    double* tab = slot + lay.tabA; double* otab = slot + lay.otab;
    double* Q0 = slot + lay.Q0; double* QO0 = slot + lay.QO0; double* G = slot + lay.G;
    unsigned long long nt = (unsigned long long)NPLANE * q.cells * S;
    // ---- start / inner posteriors (calc_motif_start_position, motif_scanner.hpp:186-193)
    cta_inside_warp(m, q, tab, otab, NoConstraint(), wsm, ctr, Q0, (double*)nullptr, nt);
    double ZL = part_func(h, otab, L, S, true, true);
    cta_zero(QO0, (unsigned long long)(L + 1) * S); cta_zero(G, (unsigned long long)h.M * L);
    cta_zero(sm.en, (unsigned long long)NT); cta_zero(sm.pys, L); cta_zero(sm.pyi, L); cta_zero(sm.pye, L + 1);
    CTA_SYNC();
    double root[3];
    root[0] = h.s00 >= 0 ? exp(otab[L * S + h.s00] - ZL) : 0.;
    root[1] = h.s0M2 >= 0 ? exp(otab[L * S + h.s0M2] - ZL) : 0.;
    root[2] = h.s0M1 >= 0 ? exp(otab[L * S + h.s0M1] - ZL) : 0.;
    Counts cn; cn.G = G; cn.ENp = sm.en; cn.Pys = sm.pys; cn.Pyi = sm.pyi; cn.Pye = sm.pye; cn.n_theta = NT;
    cn.ML = h.M * L;
    double eh[2];
    cta_outside_warp<1, HOOK_SCAN_START, false>(m, q, tab, otab, Q0, nullptr, QO0, nullptr, root, NoConstraint(), cn, eh, wsm, ctr);
    CTA_SYNC();
    if (!m.p.no_prf) cta_fold_G(m, q, G, sm.en, 1);
    CTA_SYNC();
    for (int t = CTA_TID; t < NT; t += CTA_NTH) out.EN[(long long)n * NT + t] = sm.en[t];
    for (int t = CTA_TID; t < L; t += CTA_NTH) {
      double a = sm.pys[t], c = sm.pyi[t];
      sm.pys[t] = a > 0. ? log(a) : NINF;
      out.PysL[o + t] = sm.pys[t];
      out.PyiL[o + t] = c > 0. ? log(c) : NINF;
    }
    CTA_SYNC();
    int* ish = (int*)(sm.red + 48);
    if (CTA_TID == 0) {
      ish[0] = last_max_index(sm.pys, L);
      double ex = 0.;  // exp(sumL(PysL)), motif_scanner.hpp:246
      for (int t = 0; t < L; ++t) if (sm.pys[t] > NINF) ex += exp(sm.pys[t]);
      out.exist[n] = ex;
      out.Ys[n] = ish[0];
      if (out.ZL) out.ZL[n] = ZL;
    }
    CTA_SYNC();
    int Ys = ish[0];
    // ---- end posteriors under the fixed start (calc_motif_end_position, :195-202)
    StartConstraint sc; sc.ys = Ys;
    cta_inside_warp(m, q, tab, otab, sc, wsm, ctr, Q0, (double*)nullptr, nt);
    double ZeL = part_func(h, otab, L, S, true, true);
    cta_zero(QO0, (unsigned long long)(L + 1) * S);
    CTA_SYNC();
    bool zfin = ZeL > NINF && ZeL < -NINF;
    root[0] = (zfin && h.s00 >= 0) ? exp(otab[L * S + h.s00] - ZeL) : 0.;
    root[1] = (zfin && h.s0M2 >= 0) ? exp(otab[L * S + h.s0M2] - ZeL) : 0.;
    root[2] = (zfin && h.s0M1 >= 0) ? exp(otab[L * S + h.s0M1] - ZeL) : 0.;
    cta_outside_warp<1, HOOK_SCAN_END, false>(m, q, tab, otab, Q0, nullptr, QO0, nullptr, root, sc, cn, eh, wsm, ctr);
    CTA_SYNC();
    for (int t = CTA_TID; t <= L; t += CTA_NTH) {
      double a = sm.pye[t];
      sm.pye[t] = a > 0. ? log(a) : NINF;
      out.PyeL[o + n + t] = sm.pye[t];
    }
    CTA_SYNC();
    if (CTA_TID == 0) { ish[1] = last_max_index(sm.pye, L + 1); out.Ye[n] = ish[1]; }
    CTA_SYNC();
    int Ye = ish[1];
    // ---- constrained Viterbi + traceback (calc_viterbi_alignment, :172-184)
    StartEndConstraint se; se.ys = Ys; se.ye = Ye;
    unsigned long long* trace = (unsigned long long*)Q0;
    unsigned long long* otrace = (unsigned long long*)QO0;
    cta_inside<true, StartEndConstraint>(m, q, tab, otab, trace, otrace, se);
    for (int t = CTA_TID; t < L; t += CTA_NTH) { out.psihat[o + t] = 0; out.rss[o + t] = ' '; }
    CTA_SYNC();
    if (CTA_TID == 0) {
      double a = h.s0M2 >= 0 ? otab[L * S + h.s0M2] : NINF;
      double c = h.s0M1 >= 0 ? otab[L * S + h.s0M1] : NINF;
      int s0 = (a < c) ? h.s0M1 : h.s0M2;
      if (s0 >= 0) trace_back(m, q, trace, otrace, n2s, (int*)(slot + lay.stack), s0, out.psihat + o, out.rss + o);
    }
    CTA_SYNC();
  }
}

// -------------------------------------------------------------------------------------------- Viterbi only
// Constrained Viterbi + traceback (calc_viterbi_alignment, motif_scanner.hpp:172-184) for sequences whose posteriors
// (hence Ys / Ye) and base-pair masks were produced by the linear-space kernels: masks are read from the header of the
// sequence's linear-space slot, Ys / Ye from the result arrays.
struct ExtMasks {
  const double* scratch;           // linear-space slots of the chunk
  unsigned long long stride;       // doubles per slot
  unsigned long long masks_off;    // offset of the bp / lf bit rows in a slot (doubles)
  int mask_words;                  // words per mask
  int base, count;                 // chunk = sequences order[base .. base+count), slot k <-> base+k
};
#ifdef RELEM_HOST_EMU
#define RELEM_VIT_KERNEL inline void
#else
#ifndef RELEM_VIT_MINB
#define RELEM_VIT_MINB (512 / RELEM_VIT_THREADS)
#endif
#define RELEM_VIT_KERNEL __global__ void __launch_bounds__(RELEM_VIT_THREADS, RELEM_VIT_MINB)
#endif
RELEM_VIT_KERNEL relem_viterbi_kernel(ModelView m, BatchView b, SlotLayout lay, double* scratch, int* queue, const int* n2s,
                                      ScanOut out, ExtMasks em, const unsigned char* flag RELEM_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const int S = m.h.S, R = lay.vit_R;
  const DevHMM& h = m.h;
  int* claim_sh = (int*)(smem_raw + lay.sm_vit_claim);
  VitSeq* seqs = (VitSeq*)(smem_raw + lay.sm_vit_ctx);
  VitWarp vw = vit_warp_carve(smem_raw + lay.sm_vit_warp + warp_id() * lay.vit_warp_bytes, S, lay.vit_n_max);
  // background states of cells outside the motif region (StartEndConstraint, dp_enum.cuh)
  int s_bgM = -1;
  for (int s = 0; s < S; ++s)
    if (ld_ro(h.st_l + s) == h.M - 1 && ld_ro(h.st_r + s) == h.M - 1) s_bgM = s;
  const int mask_bytes = ((lay.Lmax + 1) * lay.mw * 4 + 15) & ~15, row_bytes = (lay.Lmax + 2 + 15) & ~15;
  for (;;) {
    // ---- claim the next R sequences of the chunk
    if (CTA_TID == 0) {
#ifdef RELEM_HOST_EMU
      *claim_sh = *queue; *queue += R;
#else
      *claim_sh = atomicAdd(queue, R);
#endif
    }
    CTA_SYNC();
    const int q0 = *claim_sh;
    CTA_SYNC();
    if (q0 >= em.count) break;
    const int nr = em.count - q0 < R ? em.count - q0 : R;
    int Wmax = 0;
    // ---- set-up of every sequence of the batch (all threads)
    for (int r = 0; r < nr; ++r) {
      VitSeq& z = seqs[r];
      const int qi = q0 + r, n = b.order[em.base + qi];
      unsigned char* blk = smem_raw + lay.sm_vit_seq + r * lay.vit_seq_bytes;
      unsigned char* x = blk;
      signed char* sp3 = (signed char*)(blk + row_bytes);
      signed char* sp4 = (signed char*)(blk + 2 * row_bytes);
      signed char* sp6 = (signed char*)(blk + 3 * row_bytes);
      unsigned* bp = (unsigned*)(blk + 4 * row_bytes);
      unsigned* lf = (unsigned*)(blk + 4 * row_bytes + mask_bytes);
      double* slot = scratch + ((unsigned long long)RELEM_BLOCK_IDX * R + r) * lay.stride;
      const long long o = b.off[n];
      const int L = (int)(b.off[n + 1] - o);
      const int W = L < m.en.max_span ? L : m.en.max_span;
      const int C = W - 7 < m.en.max_iloop ? W - 7 : m.en.max_iloop;
      // the four masks the linear-space passes left in the slot header: pairs and left-ends by left end (copied to
      // shared memory: every gate test reads them) and by right end (read in place: candidate scans only)
      const unsigned* g = (const unsigned*)(em.scratch + (unsigned long long)qi * em.stride + em.masks_off);
      if (CTA_TID == 0) {
        // left the fp64 range in the linear-space pass: the full log-space kernel redoes it
        z.n = flag[n] ? -1 : n;
        z.o = o;
        SeqView& q = z.q;
        q.S = S; q.L = L; q.W = W; q.C = C; q.W1 = W + 1; q.cells = (unsigned)(L + 1) * (unsigned)(W + 1);
        q.mw = lay.mw; q.min_pair = 5; q.min_multi = 10;
        q.x = x; q.bp = bp; q.lf = lf; q.sp3 = sp3; q.sp4 = sp4; q.sp6 = sp6;
        q.ws = b.ws + o;
        q.bpr = g + 2 * (size_t)em.mask_words; q.lfr = g + 3 * (size_t)em.mask_words;
        q.emit0 = slot + lay.emit0; q.emitT = slot + lay.emitT;
        z.tab = slot + lay.tabA; z.otab = slot + lay.otab; z.stack = (int*)(slot + lay.stack);
        z.se.ys = out.Ys[n]; z.se.ye = out.Ye[n];
        z.rg.ys = z.se.ys; z.rg.ye = z.se.ye; z.rg.s_bg0 = h.s00; z.rg.s_bgM = s_bgM;
        z.rg.on = h.s00 >= 0 && s_bgM >= 0 && z.se.ys >= 0 && z.se.ye >= z.se.ys;
      }
      if (flag[n]) continue;
      if (W > Wmax) Wmax = W;
      for (int t = CTA_TID; t < L; t += CTA_NTH) { x[t] = b.seq[o + t]; out.psihat[o + t] = 0; out.rss[o + t] = ' '; }
      if (CTA_TID == 0) x[L] = 0;
      for (int t = CTA_TID; t < (L + 1) * lay.mw; t += CTA_NTH) { bp[t] = g[t]; lf[t] = g[em.mask_words + t]; }
    }
    CTA_SYNC();
    for (int r = 0; r < nr; ++r) {
      const VitSeq& z = seqs[r];
      if (z.n < 0) continue;
      cta_special_hairpins(m.en, z.q.x, z.q.L, (signed char*)z.q.sp3, (signed char*)z.q.sp4, (signed char*)z.q.sp6);
      cta_emit_tables(m, z.q, (double*)z.q.emit0, (double*)z.q.emitT);
    }
    CTA_SYNC();
    // ---- band sweep of the whole batch, then exterior row + traceback of sequence r on warp r
    cta_viterbi_band<StartEndConstraint>(m, seqs, nr, Wmax, vw);
    for (int r = warp_id(); r < nr; r += n_warps()) {
      const VitSeq& z = seqs[r];
      if (z.n < 0) continue;
      warp_viterbi_exterior<StartEndConstraint>(m, z);
      if (lane_id() == 0) {
        const int L = z.q.L;
        double a = h.s0M2 >= 0 ? z.otab[L * S + h.s0M2] : NINF;
        double c = h.s0M1 >= 0 ? z.otab[L * S + h.s0M1] : NINF;
        int s0 = (a < c) ? h.s0M1 : h.s0M2;
        if (s0 >= 0) vit_trace_back(m, z.q, z.tab, z.otab, z.se, z.rg, n2s, z.stack, s0, out.psihat + z.o, out.rss + z.o);
      }
      w_sync();
    }
    CTA_SYNC();
  }
}

}  // namespace dp
}  // namespace relem
#endif
