// relem_lin.cu -- persistent E-step kernel of the scaled linear-space path (dp_lin.cuh) and its launcher.
//
// One CTA per resident sequence ("slot"), sequences claimed longest-first from an atomic queue.  Per sequence:
//   1. set-up: bases, exp(position weights), special-hairpin hits, canonical-pair masks in both orientations;
//   2. energy-only inside + outside in gather form -> base-pair posteriors -> bp_ok / left_bp_ok masks
//      (EnergyModel::fill_bpp_tables, energy_model.hpp:211-266);
//   3. coupled inside (wavefront over the span, warp per cell), exterior row, partition functions;
//   4. coupled outside in gather form with the expected counts (RNAelemTrainDP, motif_trainer.hpp:204-245).
// Built by nvcc for sm_100a (FMA contraction allowed: nothing here has to be bit-exact) and, with
// -DRELEM_HOST_EMU under g++, into the single-threaded debug emulation used by the CPU tests.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dp_lin.cuh"
#include "dp_pass.cuh"
#include "lin_api.hpp"

#ifndef RELEM_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace relem {
namespace lin {
using namespace relem::dp;

#define LIN_THREADS 128

struct LinLayout {
  unsigned long long stride;  // doubles per slot
  unsigned long long aP, aE, aM, a1, a2, aLl, aLr, aO;
  unsigned long long bP, bEl, bEr, bM, bBl, bBr, b2, bL, bO, bch, boch;
  unsigned long long kP, kE, kM, k1, k2, kO, kbP, kbE, kbM, kbBl, kbBr, kb2, kbO;
  unsigned long long hdr, masks;  // per-slot header: Z^tt, Z^tf, Z^ft, bad; filtered bp / lf bit rows
  int Lmax, Wmax, mw;
  int sm_ctx, sm_x, sm_sp3, sm_sp4, sm_sp6, sm_bp, sm_lf, sm_bpr, sm_lfr, sm_wsf, sm_k0pow, sm_red, sm_ctr, sm_total_k0;
  int sm_en, sm_eh, sm_pcnt, sm_warp, warp_bytes_in, warp_bytes_out, sm_total_in, sm_total_out;
};

static LinLayout make_lin_layout(int Lmax, int max_span, const LinHMM& h, int n_theta, int nch, int nwarps) {
  LinLayout lay;
  std::memset(&lay, 0, sizeof(lay));
  int Wmax = Lmax < max_span ? Lmax : max_span;
  lay.Lmax = Lmax; lay.Wmax = Wmax; lay.mw = (Wmax + 1 + 31) / 32;
  unsigned long long cells = (unsigned long long)(Lmax + 1) * (Wmax + 1);
  unsigned long long band = cells * h.S, ext = (unsigned long long)(Lmax + 1) * h.S;
  unsigned long long o = 0;
  auto take = [&](unsigned long long n) { unsigned long long r = o; o += (n + 1) & ~1ull; return r; };
  lay.aP = take(band); lay.aE = take(band); lay.aM = take(band); lay.a1 = take(band); lay.a2 = take(band);
  lay.aLl = take(band); lay.aLr = take(band); lay.aO = take(ext);
  lay.bP = take(band * nch); lay.bEl = take(band * nch); lay.bEr = take(band * nch); lay.bM = take(band * nch);
  lay.bBl = take(band * nch); lay.bBr = take(band * nch); lay.b2 = take(band * nch); lay.bL = take(band * nch);
  lay.bO = take(ext * nch);
  lay.bch = band; lay.boch = ext;
  lay.kP = take(cells); lay.kE = take(cells); lay.kM = take(cells); lay.k1 = take(cells); lay.k2 = take(cells);
  lay.kO = take(Lmax + 1);
  lay.kbP = take(cells); lay.kbE = take(cells); lay.kbM = take(cells); lay.kbBl = take(cells); lay.kbBr = take(cells);
  lay.kb2 = take(cells); lay.kbO = take(Lmax + 1);
  lay.hdr = take(8);
  lay.masks = take((unsigned long long)(Lmax + 2) * lay.mw);  // 2 masks x 4-byte words
  lay.stride = o;
  int b = 0;
  auto sm = [&](int bytes) { int r = b; b += (bytes + 15) & ~15; return r; };
  lay.sm_ctx = sm((int)sizeof(LinCtx));
  lay.sm_x = sm(Lmax + 2);
  lay.sm_sp3 = sm(Lmax + 1); lay.sm_sp4 = sm(Lmax + 1); lay.sm_sp6 = sm(Lmax + 1);
  int mask_bytes = (Lmax + 2) * lay.mw * 4;
  lay.sm_bp = sm(mask_bytes); lay.sm_lf = sm(mask_bytes); lay.sm_bpr = sm(mask_bytes); lay.sm_lfr = sm(mask_bytes);
  lay.sm_wsf = sm((Lmax + 1) * 8);
  lay.sm_k0pow = sm((Wmax + 3) * 8);
  lay.sm_red = sm(64 * 8);
  lay.sm_ctr = sm(16);
  lay.sm_total_k0 = b;
  lay.sm_en = sm(nch * n_theta * 8 + 8);
  lay.sm_eh = sm(8 * 8);
  lay.sm_pcnt = sm(nch * h.n_pair * 25 * 8 + 8);
  lay.sm_warp = b;
  lay.warp_bytes_in = warp_lin_bytes(h.S, Wmax, 1, h.n_max, 0, 0, true);
  lay.warp_bytes_out = warp_lin_bytes(h.S, Wmax, nch, h.n_max, h.n_right, h.n_left, false);
  lay.sm_total_in = lay.sm_warp + lay.warp_bytes_in * nwarps;
  lay.sm_total_out = lay.sm_warp + lay.warp_bytes_out * nwarps;
  return lay;
}

struct LinKArgs {
  BatchView b;
  int base, count;   // this launch handles sequences order[base .. base+count), CTA k <-> slot k
  LinLayout lay;
  double* scratch;
  EstepOut out;
  unsigned char* flag;
};

// bp / left masks re-indexed by the right end of the span: bit d of row j <-> (j-d, j)
RDEV void cta_right_mask(const SeqView& q, const unsigned* byleft, unsigned* byright) {
  const int L = q.L, W = q.W, mw = q.mw;
  for (int t = CTA_TID; t < (L + 1) * mw; t += CTA_NTH) {
    int j = t / mw, w = t % mw;
    unsigned bits = 0u;
    for (int b = 0; b < 32; ++b) {
      int d = w * 32 + b, i = j - d;
      if (d <= W && i >= 0 && row_bit(byleft + i * mw, d)) bits |= 1u << b;
    }
    byright[t] = bits;
  }
}

// run `cell(i)` for every cell of diagonal d, cells handed to warps through a shared counter
template <class F> RDEV void lin_diagonal(int ncell, int d, int* ctr, F cell) {
  int* c = ctr + (d & 1);
  if (CTA_TID == 0) ctr[(d + 1) & 1] = 0;
  for (;;) {
    int i = 0;
    if (lane_id() == 0) i = ctr_next(c);
    i = w_shfl(i, 0);
    if (i >= ncell) break;
    cell(i);
  }
  CTA_SYNC();
}

#ifdef RELEM_HOST_EMU
#define LIN_KERNEL(MINB) inline void
#define LIN_SMEM_ARG , unsigned char* smem_raw, int emu_block
#define LIN_BLOCK_IDX emu_block
#else
#define LIN_KERNEL(MINB) __global__ void __launch_bounds__(LIN_THREADS, MINB)
#define LIN_SMEM_ARG
#define LIN_BLOCK_IDX ((int)blockIdx.x)
#endif

RDEV bool finite_pos(double v) { return v > 0. && v < (-NINF); }

// per-sequence set-up common to the three kernels: the CTA's LinCtx (in shared memory), bases, exp(position weights),
// special hairpins.  Returns the context; masks are filled by the caller.
RDEV LinCtx& lin_setup(const LinKArgs& a, unsigned char* smem_raw, int n) {
  const LinLayout& lay = a.lay;
  LinCtx& c = *(LinCtx*)(smem_raw + lay.sm_ctx);
  unsigned char* sx = smem_raw + lay.sm_x;
  double* wsf = (double*)(smem_raw + lay.sm_wsf);
  double* k0pow = (double*)(smem_raw + lay.sm_k0pow);
  const long long o = a.b.off[n];
  const int L = (int)(a.b.off[n + 1] - o);
  if (CTA_TID == 0) {
    const int W = L < LC.en.max_span ? L : LC.en.max_span;
    const int C = W - 7 < LC.en.max_iloop ? W - 7 : LC.en.max_iloop;
    SeqView& q = c.q;
    q.L = L; q.W = W; q.C = C; q.W1 = W + 1; q.S = LC.h.S; q.cells = (unsigned)(L + 1) * (unsigned)(W + 1);
    q.mw = lay.mw; q.min_pair = 5; q.min_multi = 10;
    q.x = sx;
    q.bp = (unsigned*)(smem_raw + lay.sm_bp); q.lf = (unsigned*)(smem_raw + lay.sm_lf);
    q.sp3 = (signed char*)(smem_raw + lay.sm_sp3); q.sp4 = (signed char*)(smem_raw + lay.sm_sp4);
    q.sp6 = (signed char*)(smem_raw + lay.sm_sp6);
    q.ws = a.b.ws + o; q.emit0 = nullptr; q.emitT = nullptr;
    c.bpr = (unsigned*)(smem_raw + lay.sm_bpr); c.lfr = (unsigned*)(smem_raw + lay.sm_lfr);
    c.wsf = wsf; c.k0pow = k0pow;
    c.Ceff = C < 30 ? C : 30;
    sx[L] = 0; sx[L + 1] = 0;
    int* ctr = (int*)(smem_raw + lay.sm_ctr);
    ctr[0] = 0; ctr[1] = 0;
  }
  for (int t = CTA_TID; t < L; t += CTA_NTH) { sx[t] = a.b.seq[o + t]; wsf[t] = exp(a.b.ws[o + t]); }
  CTA_SYNC();
  for (int t = CTA_TID; t <= c.q.W + 2; t += CTA_NTH) k0pow[t] = pow(LC.k0, (double)t);
  cta_special_hairpins(LC.en, sx, L, (signed char*)(smem_raw + lay.sm_sp3), (signed char*)(smem_raw + lay.sm_sp4),
                       (signed char*)(smem_raw + lay.sm_sp6));
  CTA_SYNC();
  return c;
}
// filtered masks of this slot (written by the filter kernel) -> shared memory, both orientations
RDEV void lin_load_masks(const LinKArgs& a, unsigned char* smem_raw, const LinCtx& c, const double* slot) {
  const LinLayout& lay = a.lay;
  const unsigned* g = (const unsigned*)(slot + lay.masks);
  unsigned* bp = (unsigned*)(smem_raw + lay.sm_bp);
  unsigned* lf = (unsigned*)(smem_raw + lay.sm_lf);
  const int nw = (c.q.L + 1) * lay.mw;
  for (int t = CTA_TID; t < nw; t += CTA_NTH) { bp[t] = g[t]; lf[t] = g[nw + t]; }
  CTA_SYNC();
  cta_right_mask(c.q, bp, (unsigned*)(smem_raw + lay.sm_bpr));
  cta_right_mask(c.q, lf, (unsigned*)(smem_raw + lay.sm_lfr));
  CTA_SYNC();
}

// ------------------------------------------------------------------------------------------------ kernel A
// energy-only inside/outside -> filtered base-pair masks of every sequence of the chunk
LIN_KERNEL(8) relem_lin_filter_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const LinLayout& lay = a.lay;
  const int blk = LIN_BLOCK_IDX;
  if (blk >= a.count) return;
  const int n = a.b.order[a.base + blk];
  double* slot = a.scratch + (unsigned long long)blk * lay.stride;
  LinCtx& c = lin_setup(a, smem_raw, n);
  const SeqView& q = c.q;
  unsigned* bp = (unsigned*)(smem_raw + lay.sm_bp);
  unsigned* lf = (unsigned*)(smem_raw + lay.sm_lf);
  unsigned* bpr = (unsigned*)(smem_raw + lay.sm_bpr);
  unsigned* lfr = (unsigned*)(smem_raw + lay.sm_lfr);
  double* red = (double*)(smem_raw + lay.sm_red);
  int* ctr = (int*)(smem_raw + lay.sm_ctr);
  const int L = q.L, W = q.W;
  cta_canonical_mask(q, bp);
  CTA_SYNC();
  cta_left_mask(q, bp, lf);
  const int total = cta_count_bits(bp, (L + 1) * lay.mw, (int*)red);
  int nbp = total;
  bool bad = false;
  if (LC.en.filter) {
    cta_right_mask(q, bp, bpr);
    cta_right_mask(q, lf, lfr);
    CTA_SYNC();
    K0Tabs t0;
    t0.P = slot + lay.kP; t0.E = slot + lay.kE; t0.M = slot + lay.kM; t0.o1 = slot + lay.k1; t0.o2 = slot + lay.k2;
    t0.O = slot + lay.kO; t0.bP = slot + lay.kbP; t0.bE = slot + lay.kbE; t0.bM = slot + lay.kbM;
    t0.bBl = slot + lay.kbBl; t0.bBr = slot + lay.kbBr; t0.b2 = slot + lay.kb2; t0.bO = slot + lay.kbO;
    for (int d = 3; d <= W; ++d) lin_diagonal(L + 1 - d, d, ctr, [&](int i) { k0_inside_cell(c, t0, i, d); });
    if (CTA_TID == 0) { ctr[0] = 0; ctr[1] = 0; }
    if (warp_id() == 0) k0_inside_ext(c, t0);
    CTA_SYNC();
    const double Z0 = ld_cg(t0.O + L);
    bad = !finite_pos(Z0);
    if (!bad) {
      if (warp_id() == 0) k0_outside_ext(c, t0, 1. / Z0);
      CTA_SYNC();
      for (int d = W; d >= 3; --d) lin_diagonal(L + 1 - d, d, ctr, [&](int i) { k0_outside_cell(c, t0, i, d); });
      // keep pairs with ln BPP >= ln min_bpp (energy_model.hpp:257-261)
      for (int t = CTA_TID; t < (L + 1) * lay.mw; t += CTA_NTH) {
        int i = t / lay.mw, ww = t % lay.mw;
        unsigned in = bp[t], outb = 0u;
        for (int bb = 0; bb < 32; ++bb) {
          if (!((in >> bb) & 1u)) continue;
          int d = ww * 32 + bb;
          double post = ld_cg(t0.P + kidx(q, i + d, d)) * ld_cg(t0.bP + kidx(q, i, d));
          double ln = post > 0. ? log(post) : NINF;
          if (LC.en.min_lnbpp <= ln) outb |= 1u << bb;
        }
        bp[t] = outb;
      }
      CTA_SYNC();
      cta_left_mask(q, bp, lf);
      nbp = cta_count_bits(bp, (L + 1) * lay.mw, (int*)red);
    }
  }
  CTA_SYNC();
  unsigned* g = (unsigned*)(slot + lay.masks);
  const int nw = (L + 1) * lay.mw;
  for (int t = CTA_TID; t < nw; t += CTA_NTH) { g[t] = bp[t]; g[nw + t] = lf[t]; }
  if (CTA_TID == 0) {
    slot[lay.hdr + 3] = bad ? 1. : 0.;
    a.out.bpp_eff[n] = (double)nbp / (double)total;
    if (bad) a.flag[n] = 1;
  }
}

RDEV CTabs lin_tabs(const LinLayout& lay, double* slot) {
  CTabs t;
  t.aP = slot + lay.aP; t.aE = slot + lay.aE; t.aM = slot + lay.aM; t.a1 = slot + lay.a1; t.a2 = slot + lay.a2;
  t.aLl = slot + lay.aLl; t.aLr = slot + lay.aLr; t.aO = slot + lay.aO;
  t.bP = slot + lay.bP; t.bEl = slot + lay.bEl; t.bEr = slot + lay.bEr; t.bM = slot + lay.bM; t.bBl = slot + lay.bBl;
  t.bBr = slot + lay.bBr; t.b2 = slot + lay.b2; t.bL = slot + lay.bL; t.bO = slot + lay.bO;
  t.bch = (unsigned)lay.bch; t.boch = (unsigned)lay.boch;
  return t;
}

// ------------------------------------------------------------------------------------------------ kernel B
// coupled inside pass + partition functions
LIN_KERNEL(5) relem_lin_inside_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const LinLayout& lay = a.lay;
  const int blk = LIN_BLOCK_IDX;
  if (blk >= a.count) return;
  const int n = a.b.order[a.base + blk];
  double* slot = a.scratch + (unsigned long long)blk * lay.stride;
  if (slot[lay.hdr + 3] != 0.) return;
  LinCtx& c = lin_setup(a, smem_raw, n);
  lin_load_masks(a, smem_raw, c, slot);
  const SeqView& q = c.q;
  const LinHMM& h = LC.h;
  const int L = q.L, W = q.W, S = q.S;
  int* ctr = (int*)(smem_raw + lay.sm_ctr);
  WarpLin w = warp_lin_carve(smem_raw + lay.sm_warp + warp_id() * lay.warp_bytes_in, S, lay.Wmax, 1, h.n_max, 0, 0, true);
  CTabs t = lin_tabs(lay, slot);
  for (int d = 0; d <= W; ++d) { lin_inside_diag(c, t, d, w); CTA_SYNC(); }
  if (warp_id() == 0) lin_inside_ext(c, t, w);
  CTA_SYNC();
  if (CTA_TID == 0) {
    const double r00 = h.s00 >= 0 ? ld_cg(t.aO + (unsigned)L * S + h.s00) : 0.;
    const double rM2 = h.s0M2 >= 0 ? ld_cg(t.aO + (unsigned)L * S + h.s0M2) : 0.;
    const double rM1 = h.s0M1 >= 0 ? ld_cg(t.aO + (unsigned)L * S + h.s0M1) : 0.;
    const double Ztt = r00 + (rM2 + rM1), Ztf = rM2 + rM1, Zft = r00;
    const int kind = a.b.kind[n];
    // every partition function the trainer tests must be representable; otherwise the log-space path decides
    const bool bad = !finite_pos(Ztt) || (kind != 2 && !finite_pos(Ztf)) || !(Zft >= 0. && Zft < (-NINF));
    slot[lay.hdr + 0] = Ztt; slot[lay.hdr + 1] = Ztf; slot[lay.hdr + 2] = Zft;
    if (bad) { slot[lay.hdr + 3] = 1.; a.flag[n] = 1; }
    else {
      const double shift = -(double)L * LC.p.ln_kappa;  // ln Z = ln Z^ - L ln kappa
      a.out.Z[n * 3 + 0] = log(Ztt) + shift;
      a.out.Z[n * 3 + 1] = Ztf > 0. ? log(Ztf) + shift : NINF;
      a.out.Z[n * 3 + 2] = Zft > 0. ? log(Zft) + shift : NINF;
      a.out.skipped[n] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ kernel C
// coupled outside pass in gather form + expected counts
template <int NCH> LIN_KERNEL(5) relem_lin_outside_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const LinLayout& lay = a.lay;
  const int blk = LIN_BLOCK_IDX;
  if (blk >= a.count) return;
  const int n = a.b.order[a.base + blk];
  double* slot = a.scratch + (unsigned long long)blk * lay.stride;
  if (slot[lay.hdr + 3] != 0.) return;
  LinCtx& c = lin_setup(a, smem_raw, n);
  lin_load_masks(a, smem_raw, c, slot);
  const SeqView& q = c.q;
  const LinHMM& h = LC.h;
  const int L = q.L, W = q.W, S = q.S, NT = LC.p.n_theta;
  int* ctr = (int*)(smem_raw + lay.sm_ctr);
  double* sen = (double*)(smem_raw + lay.sm_en);
  double* seh = (double*)(smem_raw + lay.sm_eh);
  double* pcnt = (double*)(smem_raw + lay.sm_pcnt);
  WarpLin w = warp_lin_carve(smem_raw + lay.sm_warp + warp_id() * lay.warp_bytes_out, S, lay.Wmax, NCH, h.n_max,
                             h.n_right, h.n_left, false);
  w.pcnt = pcnt;
  CTabs t = lin_tabs(lay, slot);
  const double Ztt = slot[lay.hdr + 0], Ztf = slot[lay.hdr + 1], Zft = slot[lay.hdr + 2];
  const int kind = a.b.kind[n];
  // root weights of the outside pass: channel 0 = Zo (all three roots), channel 1 = the restricted condition
  double rw[NCH][3];
  {
    double o0 = 1. / Ztt;
    double x00 = 0., xM = 0.;
    if (kind == 1) xM = 1. / Ztf;
    else x00 = Zft > 0. ? 1. / Zft : 0.;
    if (NCH == 2) {
      rw[0][0] = o0; rw[0][1] = o0; rw[0][2] = o0;
      rw[NCH - 1][0] = x00; rw[NCH - 1][1] = xM; rw[NCH - 1][2] = xM;
    } else {
      rw[0][0] = o0 - x00; rw[0][1] = o0 - xM; rw[0][2] = o0 - xM;
    }
  }
  for (int tt = CTA_TID; tt < NCH * S; tt += CTA_NTH) {
    int ch = tt / S, s = tt - ch * S;
    double v = 0.;
    if (s == h.s00) v = rw[ch][0];
    if (s == h.s0M2) v = rw[ch][1];
    if (s == h.s0M1) v = rw[ch][2];
    t.bO[ch * t.boch + (unsigned)L * S + s] = v;
  }
  for (int tt = CTA_TID; tt < NCH * NT; tt += CTA_NTH) sen[tt] = 0.;
  for (int tt = CTA_TID; tt < NCH * h.n_pair * 25; tt += CTA_NTH) pcnt[tt] = 0.;
  for (int tt = CTA_TID; tt < 8; tt += CTA_NTH) seh[tt] = 0.;
  for (int tt = lane_id(); tt < NCH * 5 * h.n_right; tt += WARP_N) w.cntR[tt] = 0.;
  for (int tt = lane_id(); tt < NCH * 5 * h.n_left; tt += WARP_N) w.cntL[tt] = 0.;
  CTA_SYNC();
  EhAcc<NCH> eh;
  for (int k = 0; k < NCH * 2; ++k) eh.v[k] = 0.;
  if (warp_id() == 0) lin_outside_ext<NCH>(c, t, w);
  CTA_SYNC();
  for (int d = W; d >= 0; --d) { lin_outside_diag<NCH>(c, t, d, w, eh); CTA_SYNC(); }
  // ---- fold the per-entry emission sums into theta-shaped counts
  w_sync();
  if (!LC.p.no_prf) {
    for (int tt = lane_id(); tt < NCH * 5 * h.n_right; tt += WARP_N) {
      int ch = tt / (5 * h.n_right), r = tt - ch * 5 * h.n_right;
      int idx = ld_ro(h.r_en + r);
      double v = w.cntR[tt];
      if (idx >= 0 && v != 0.) sm_add(sen + ch * NT + idx, v);
    }
    for (int tt = lane_id(); tt < NCH * 5 * h.n_left; tt += WARP_N) {
      int ch = tt / (5 * h.n_left), r = tt - ch * 5 * h.n_left;
      int idx = ld_ro(h.l_en + r);
      double v = w.cntL[tt];
      if (idx >= 0 && v != 0.) sm_add(sen + ch * NT + idx, v);
    }
  }
  for (int k = 0; k < NCH * 2; ++k) {
    double v = w_sum(eh.v[k]);
    if (lane_id() == 0 && v != 0.) sm_add(seh + k, v);
  }
  CTA_SYNC();
  if (!LC.p.no_prf) {
    for (int tt = CTA_TID; tt < NCH * h.n_pair * 25; tt += CTA_NTH) {
      int ch = tt / (h.n_pair * 25), r = tt - ch * h.n_pair * 25;
      double v = pcnt[tt];
      if (v == 0.) continue;
      int i1 = ld_ro(h.p_en1 + r), i2 = ld_ro(h.p_en2 + r);
      if (i1 >= 0) sm_add(sen + ch * NT + i1, v);
      if (i2 >= 0) sm_add(sen + ch * NT + i2, v);
    }
  }
  CTA_SYNC();
  // ---- results; non-finite counts mean the scaled tables overflowed somewhere: let the log-space path redo it
  bool okv = true;
  for (int tt = 0; tt < NCH * NT; ++tt) okv = okv && (sen[tt] - sen[tt] == 0.);
  for (int k = 0; k < NCH * 2; ++k) okv = okv && (seh[k] - seh[k] == 0.);
  if (!okv) {
    if (CTA_TID == 0) a.flag[n] = 1;
    return;
  }
  for (int tt = CTA_TID; tt < NT; tt += CTA_NTH) {
    a.out.ENo[(long long)n * NT + tt] = sen[tt];
    a.out.ENx[(long long)n * NT + tt] = NCH == 2 ? sen[(NCH - 1) * NT + tt] : 0.;
  }
  if (CTA_TID == 0) {
    a.out.EH[n * 4 + 0] = seh[0]; a.out.EH[n * 4 + 1] = seh[1];
    a.out.EH[n * 4 + 2] = NCH == 2 ? seh[(NCH - 1) * 2] : 0.;
    a.out.EH[n * 4 + 3] = NCH == 2 ? seh[(NCH - 1) * 2 + 1] : 0.;
  }
}

// ------------------------------------------------------------------------------------------------ launcher
struct LinState {
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
};

LinState* lin_state_create() { return new LinState(); }
void lin_state_destroy(LinState* s) {
  if (!s) return;
#ifdef RELEM_HOST_EMU
  std::free(s->scratch);
#else
  if (s->scratch) cudaFree(s->scratch);
#endif
  delete s;
}

int lin_estep_launch(LinState* st, const LinLaunch& in, float* kernel_ms, int* launches, std::string& err) {
  if (kernel_ms) *kernel_ms = 0.f;
  if (launches) *launches = 0;
  const int nwarps = LIN_THREADS / 32;
  LinConst hc;
  hc.h = in.h; hc.p = in.p; hc.en = in.en; hc.el = in.el; hc.k0 = in.kappa0; hc.k0sq = in.kappa0 * in.kappa0;
  LinKArgs a;
  a.b = in.b; a.out = in.out; a.flag = in.flag;
  const int nseq = in.b.nseq;
#ifdef RELEM_HOST_EMU
  LC = hc;
  a.lay = make_lin_layout(std::max(1, in.Lmax), in.max_span, in.h, in.p.n_theta, in.nch, 1);
  size_t need = (size_t)a.lay.stride * sizeof(double);
  if (need > st->scratch_bytes) {
    std::free(st->scratch);
    st->scratch = std::malloc(need);
    st->scratch_bytes = st->scratch ? need : 0;
  }
  if (!st->scratch) { err = "scratch allocation failed"; return 3; }
  a.scratch = (double*)st->scratch;
  std::vector<unsigned char> smem(std::max(a.lay.sm_total_in, a.lay.sm_total_out) + 64);
  for (int k = 0; k < nseq; ++k) {
    // poison: the gather passes must never read an entry they did not write
    { double* p = (double*)st->scratch; for (size_t z = 0; z < need / 8; ++z) p[z] = std::nan(""); }
    a.base = k; a.count = 1;
    relem_lin_filter_kernel(a, smem.data(), 0);
    relem_lin_inside_kernel(a, smem.data(), 0);
    if (in.nch == 2) relem_lin_outside_kernel<2>(a, smem.data(), 0);
    else relem_lin_outside_kernel<1>(a, smem.data(), 0);
  }
  if (launches) *launches = 3 * nseq;
  (void)nwarps;
  return 0;
#else
  a.lay = make_lin_layout(std::max(1, in.Lmax), in.max_span, in.h, in.p.n_theta, in.nch, nwarps);
  if (a.lay.sm_total_out > 227 * 1024) { err = "sequence too long for the linear-space kernel's shared memory"; return 1; }
  const void* kout = in.nch == 2 ? (const void*)relem_lin_outside_kernel<2> : (const void*)relem_lin_outside_kernel<1>;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute((const void*)relem_lin_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, a.lay.sm_total_k0)) != cudaSuccess ||
      (e = cudaFuncSetAttribute((const void*)relem_lin_inside_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, a.lay.sm_total_in)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(kout, cudaFuncAttributeMaxDynamicSharedMemorySize, a.lay.sm_total_out)) != cudaSuccess) {
    err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
    return 2;
  }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  size_t per = (size_t)a.lay.stride * sizeof(double);
  long long by_mem = (long long)(((double)(free_b + st->scratch_bytes) * 0.70) / (double)per);
  long long nslots = std::min<long long>(std::min<long long>(nseq, (long long)in.sm_count * 32), by_mem);
  if (in.max_slots > 0) nslots = std::min<long long>(nslots, in.max_slots);
  if (nslots < 1) { err = "not enough device memory for one sequence slot"; return 3; }
  size_t need = (size_t)nslots * per;
  if (need > st->scratch_bytes) {
    if (st->scratch) cudaFree(st->scratch);
    st->scratch = nullptr; st->scratch_bytes = 0;
    if (cudaMalloc(&st->scratch, need) != cudaSuccess) { err = "scratch allocation failed"; return 3; }
    st->scratch_bytes = need;
  }
  cudaStream_t stream = (cudaStream_t)in.stream;
  e = cudaMemcpyToSymbolAsync(LC, &hc, sizeof(LinConst), 0, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) { err = std::string("constant upload: ") + cudaGetErrorString(e); return 2; }
  a.scratch = (double*)st->scratch;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  int nl = 0;
  for (int base = 0; base < nseq; base += (int)nslots) {
    a.base = base; a.count = std::min<int>((int)nslots, nseq - base);
    relem_lin_filter_kernel<<<a.count, LIN_THREADS, a.lay.sm_total_k0, stream>>>(a);
    relem_lin_inside_kernel<<<a.count, LIN_THREADS, a.lay.sm_total_in, stream>>>(a);
    if (in.nch == 2) relem_lin_outside_kernel<2><<<a.count, LIN_THREADS, a.lay.sm_total_out, stream>>>(a);
    else relem_lin_outside_kernel<1><<<a.count, LIN_THREADS, a.lay.sm_total_out, stream>>>(a);
    nl += 3;
  }
  e = cudaGetLastError();
  cudaEventRecord(e1, stream);
  if (e != cudaSuccess) { err = std::string("linear-space kernel launch: ") + cudaGetErrorString(e); return 2; }
  e = cudaEventSynchronize(e1);
  if (e != cudaSuccess) { err = std::string("linear-space kernels: ") + cudaGetErrorString(e); return 2; }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (kernel_ms) *kernel_ms = ms;
  if (launches) *launches = nl;
  return 0;
#endif
}

}  // namespace lin
}  // namespace relem
