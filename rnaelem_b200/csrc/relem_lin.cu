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

#define LIN_THREADS 256

struct LinLayout {
  unsigned long long stride;  // doubles per slot
  unsigned long long aP, aE, aM, a1, a2, aLl, aLr, aO;
  unsigned long long bP, bEl, bEr, bM, bBl, bBr, b2, bL, bO, bch, boch;
  unsigned long long kP, kE, kM, k1, k2, kO, kbP, kbE, kbM, kbBl, kbBr, kb2, kbO;
  int Lmax, Wmax, mw;
  int sm_x, sm_sp3, sm_sp4, sm_sp6, sm_bp, sm_lf, sm_bpr, sm_lfr, sm_wsf, sm_k0pow, sm_en, sm_eh, sm_pcnt, sm_red, sm_ctr,
      sm_warp, warp_bytes, sm_total;
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
  lay.stride = o;
  int b = 0;
  auto sm = [&](int bytes) { int r = b; b += (bytes + 15) & ~15; return r; };
  lay.sm_x = sm(Lmax + 2);
  lay.sm_sp3 = sm(Lmax + 1); lay.sm_sp4 = sm(Lmax + 1); lay.sm_sp6 = sm(Lmax + 1);
  int mask_bytes = (Lmax + 2) * lay.mw * 4;
  lay.sm_bp = sm(mask_bytes); lay.sm_lf = sm(mask_bytes); lay.sm_bpr = sm(mask_bytes); lay.sm_lfr = sm(mask_bytes);
  lay.sm_wsf = sm((Lmax + 1) * 8);
  lay.sm_k0pow = sm((Wmax + 3) * 8);
  lay.sm_en = sm(nch * n_theta * 8 + 8);
  lay.sm_eh = sm(8 * 8);
  lay.sm_pcnt = sm(nch * h.n_pair * 25 * 8 + 8);
  lay.sm_red = sm(64 * 8);
  lay.sm_ctr = sm(16);
  lay.warp_bytes = warp_lin_bytes(h.S, Wmax, nch, h.n_max, h.n_right, h.n_left);
  lay.sm_warp = sm(lay.warp_bytes * nwarps);
  lay.sm_total = b;
  return lay;
}

struct LinKArgs {
  LinHMM h;
  LinParams p;
  DevEnergy en, el;
  double kappa0;
  BatchView b;
  LinLayout lay;
  double* scratch;
  int* queue;
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

RDEV int lin_claim(int* queue, int* sh) {
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
#define LIN_KERNEL inline void
#define LIN_SMEM_ARG , unsigned char* smem_raw
#define LIN_BLOCK_IDX 0
#else
#define LIN_KERNEL __global__ void __launch_bounds__(LIN_THREADS)
#define LIN_SMEM_ARG
#define LIN_BLOCK_IDX ((int)blockIdx.x)
#endif

RDEV bool finite_pos(double v) { return v > 0. && v < (-NINF); }

template <int NCH> LIN_KERNEL relem_estep_lin_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const LinLayout& lay = a.lay;
  double* slot = a.scratch + (unsigned long long)LIN_BLOCK_IDX * lay.stride;
  unsigned char* sx = smem_raw + lay.sm_x;
  signed char* sp3 = (signed char*)(smem_raw + lay.sm_sp3);
  signed char* sp4 = (signed char*)(smem_raw + lay.sm_sp4);
  signed char* sp6 = (signed char*)(smem_raw + lay.sm_sp6);
  unsigned* bp = (unsigned*)(smem_raw + lay.sm_bp);
  unsigned* lf = (unsigned*)(smem_raw + lay.sm_lf);
  unsigned* bpr = (unsigned*)(smem_raw + lay.sm_bpr);
  unsigned* lfr = (unsigned*)(smem_raw + lay.sm_lfr);
  double* wsf = (double*)(smem_raw + lay.sm_wsf);
  double* k0pow = (double*)(smem_raw + lay.sm_k0pow);
  double* sen = (double*)(smem_raw + lay.sm_en);
  double* seh = (double*)(smem_raw + lay.sm_eh);
  double* pcnt = (double*)(smem_raw + lay.sm_pcnt);
  double* red = (double*)(smem_raw + lay.sm_red);
  int* ctr = (int*)(smem_raw + lay.sm_ctr);
  const int S = a.h.S, NT = a.p.n_theta;
  WarpLin w = warp_lin_carve(smem_raw + lay.sm_warp + warp_id() * lay.warp_bytes, S, lay.Wmax, NCH, a.h.n_max,
                             a.h.n_right, a.h.n_left);
  w.pcnt = pcnt;
  if (CTA_TID == 0) { ctr[0] = 0; ctr[1] = 0; }
  CTA_SYNC();
  LinCtx c;
  c.h = a.h; c.p = a.p; c.en = a.en; c.el = a.el;
  c.bpr = bpr; c.lfr = lfr; c.wsf = wsf; c.k0pow = k0pow;
  c.k0 = a.kappa0; c.k0sq = a.kappa0 * a.kappa0;
  const LinHMM& h = c.h;
  for (;;) {
    int qi = lin_claim(a.queue, (int*)(red + 40));
    if (qi >= a.b.nseq) break;
    const int n = a.b.order[qi];
    const long long o = a.b.off[n];
    const int L = (int)(a.b.off[n + 1] - o);
    const int W = L < a.en.max_span ? L : a.en.max_span;
    const int C = W - 7 < a.en.max_iloop ? W - 7 : a.en.max_iloop;
    SeqView& q = c.q;
    q.L = L; q.W = W; q.C = C; q.W1 = W + 1; q.S = S; q.cells = (unsigned)(L + 1) * (unsigned)(W + 1);
    q.mw = lay.mw; q.min_pair = 5; q.min_multi = 10;
    q.x = sx; q.bp = bp; q.lf = lf; q.sp3 = sp3; q.sp4 = sp4; q.sp6 = sp6;
    q.ws = a.b.ws + o; q.emit0 = nullptr; q.emitT = nullptr;
    c.Ceff = C < 30 ? C : 30;
    for (int t = CTA_TID; t < L; t += CTA_NTH) { sx[t] = a.b.seq[o + t]; wsf[t] = exp(a.b.ws[o + t]); }
    if (CTA_TID == 0) { sx[L] = 0; sx[L + 1] = 0; }
    for (int t = CTA_TID; t <= W + 2; t += CTA_NTH) k0pow[t] = pow(a.kappa0, (double)t);
    CTA_SYNC();
    cta_special_hairpins(a.en, sx, L, sp3, sp4, sp6);
    cta_canonical_mask(q, bp);
    CTA_SYNC();
    cta_left_mask(q, bp, lf);
    const int total = cta_count_bits(bp, (L + 1) * lay.mw, (int*)red);
    int nbp = total;
    bool bad = false;
    if (a.en.filter) {
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
      const double Z0 = t0.O[L];
      bad = !finite_pos(Z0);
      if (!bad) {
        if (warp_id() == 0) k0_outside_ext(c, t0, 1. / Z0);
        CTA_SYNC();
        for (int d = W; d >= 3; --d) lin_diagonal(L + 1 - d, d, ctr, [&](int i) { k0_outside_cell(c, t0, i, d); });
        if (CTA_TID == 0) { ctr[0] = 0; ctr[1] = 0; }
        CTA_SYNC();
        // keep pairs with ln BPP >= ln min_bpp (energy_model.hpp:257-261)
        for (int t = CTA_TID; t < (L + 1) * lay.mw; t += CTA_NTH) {
          int i = t / lay.mw, ww = t % lay.mw;
          unsigned in = bp[t], outb = 0u;
          for (int bb = 0; bb < 32; ++bb) {
            if (!((in >> bb) & 1u)) continue;
            int d = ww * 32 + bb;
            double post = ld_cg(t0.P + kidx(q, i + d, d)) * ld_cg(t0.bP + kidx(q, i, d));
            double ln = post > 0. ? log(post) : NINF;
            if (a.en.min_lnbpp <= ln) outb |= 1u << bb;
          }
          bp[t] = outb;
        }
        CTA_SYNC();
        cta_left_mask(q, bp, lf);
        nbp = cta_count_bits(bp, (L + 1) * lay.mw, (int*)red);
      }
    }
    CTA_SYNC();
    cta_right_mask(q, bp, bpr);
    cta_right_mask(q, lf, lfr);
    CTA_SYNC();
    const double eff = (double)nbp / (double)total;
    // ------------------------------------------------------------------------------- coupled passes
    CTabs t;
    t.aP = slot + lay.aP; t.aE = slot + lay.aE; t.aM = slot + lay.aM; t.a1 = slot + lay.a1; t.a2 = slot + lay.a2;
    t.aLl = slot + lay.aLl; t.aLr = slot + lay.aLr; t.aO = slot + lay.aO;
    t.bP = slot + lay.bP; t.bEl = slot + lay.bEl; t.bEr = slot + lay.bEr; t.bM = slot + lay.bM; t.bBl = slot + lay.bBl;
    t.bBr = slot + lay.bBr; t.b2 = slot + lay.b2; t.bL = slot + lay.bL; t.bO = slot + lay.bO;
    t.bch = lay.bch; t.boch = lay.boch;
    double Ztt = 0., Ztf = 0., Zft = 0.;
    const int kind = a.b.kind[n];
    if (!bad) {
      for (int d = 0; d <= W; ++d) lin_diagonal(L + 1 - d, d, ctr, [&](int i) { lin_inside_cell(c, t, i, d, w); });
      if (CTA_TID == 0) { ctr[0] = 0; ctr[1] = 0; }
      if (warp_id() == 0) lin_inside_ext(c, t, w);
      CTA_SYNC();
      const double r00 = h.s00 >= 0 ? ld_cg(t.aO + (size_t)L * S + h.s00) : 0.;
      const double rM2 = h.s0M2 >= 0 ? ld_cg(t.aO + (size_t)L * S + h.s0M2) : 0.;
      const double rM1 = h.s0M1 >= 0 ? ld_cg(t.aO + (size_t)L * S + h.s0M1) : 0.;
      Ztt = r00 + (rM2 + rM1); Ztf = rM2 + rM1; Zft = r00;
      // every partition function the trainer tests must be representable; otherwise the log-space path decides
      bad = !finite_pos(Ztt) || (kind != 2 && !finite_pos(Ztf)) || !(Zft >= 0. && Zft < (-NINF));
    }
    if (bad) {
      if (CTA_TID == 0) a.flag[n] = 1;
      CTA_SYNC();
      continue;
    }
    const double shift = -(double)L * c.p.ln_kappa;  // ln Z = ln Z^ - L ln kappa
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
      t.bO[ch * t.boch + (size_t)L * S + s] = v;
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
    for (int d = W; d >= 0; --d) lin_diagonal(L + 1 - d, d, ctr, [&](int i) { lin_outside_cell<NCH>(c, t, i, d, w, eh); });
    if (CTA_TID == 0) { ctr[0] = 0; ctr[1] = 0; }
    // ---- fold the per-entry emission sums into theta-shaped counts
    w_sync();
    if (!c.p.no_prf) {
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
    if (!c.p.no_prf) {
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
      CTA_SYNC();
      continue;
    }
    for (int tt = CTA_TID; tt < NT; tt += CTA_NTH) {
      a.out.ENo[(long long)n * NT + tt] = sen[tt];
      a.out.ENx[(long long)n * NT + tt] = NCH == 2 ? sen[(NCH - 1) * NT + tt] : 0.;
    }
    if (CTA_TID == 0) {
      a.out.Z[n * 3 + 0] = log(Ztt) + shift;
      a.out.Z[n * 3 + 1] = Ztf > 0. ? log(Ztf) + shift : NINF;
      a.out.Z[n * 3 + 2] = Zft > 0. ? log(Zft) + shift : NINF;
      a.out.bpp_eff[n] = eff;
      a.out.skipped[n] = 0;
      a.out.EH[n * 4 + 0] = seh[0]; a.out.EH[n * 4 + 1] = seh[1];
      a.out.EH[n * 4 + 2] = NCH == 2 ? seh[(NCH - 1) * 2] : 0.;
      a.out.EH[n * 4 + 3] = NCH == 2 ? seh[(NCH - 1) * 2 + 1] : 0.;
      a.flag[n] = 0;
    }
    CTA_SYNC();
  }
}

// ------------------------------------------------------------------------------------------------ launcher
struct LinState {
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* queue = nullptr;
};

LinState* lin_state_create() { return new LinState(); }
void lin_state_destroy(LinState* s) {
  if (!s) return;
#ifdef RELEM_HOST_EMU
  std::free(s->scratch); std::free(s->queue);
#else
  if (s->scratch) cudaFree(s->scratch);
  if (s->queue) cudaFree(s->queue);
#endif
  delete s;
}

int lin_estep_launch(LinState* st, const LinLaunch& in, float* kernel_ms, int* launches, std::string& err) {
  if (kernel_ms) *kernel_ms = 0.f;
  if (launches) *launches = 0;
  const int nwarps = LIN_THREADS / 32;
  LinKArgs a;
  a.h = in.h; a.p = in.p; a.en = in.en; a.el = in.el; a.kappa0 = in.kappa0; a.b = in.b; a.out = in.out; a.flag = in.flag;
#ifdef RELEM_HOST_EMU
  a.lay = make_lin_layout(std::max(1, in.Lmax), in.max_span, in.h, in.p.n_theta, in.nch, 1);
  size_t need = (size_t)a.lay.stride * sizeof(double);
  if (need > st->scratch_bytes) {
    std::free(st->scratch);
    st->scratch = std::malloc(need);
    st->scratch_bytes = st->scratch ? need : 0;
  }
  if (!st->scratch) { err = "scratch allocation failed"; return 3; }
  // poison: the gather passes must never read an entry they did not write
  { double* p = (double*)st->scratch; for (size_t k = 0; k < need / 8; ++k) p[k] = std::nan(""); }
  if (!st->queue) st->queue = std::malloc(sizeof(int));
  *(int*)st->queue = 0;
  a.scratch = (double*)st->scratch; a.queue = (int*)st->queue;
  std::vector<unsigned char> smem(a.lay.sm_total + 64);
  if (in.nch == 2) relem_estep_lin_kernel<2>(a, smem.data());
  else relem_estep_lin_kernel<1>(a, smem.data());
  if (launches) *launches = 1;
  (void)nwarps;
  return 0;
#else
  a.lay = make_lin_layout(std::max(1, in.Lmax), in.max_span, in.h, in.p.n_theta, in.nch, nwarps);
  if (a.lay.sm_total > 227 * 1024) { err = "sequence too long for the linear-space kernel's shared memory"; return 1; }
  const void* kern = in.nch == 2 ? (const void*)relem_estep_lin_kernel<2> : (const void*)relem_estep_lin_kernel<1>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.lay.sm_total);
  if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 2; }
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LIN_THREADS, a.lay.sm_total);
  if (e != cudaSuccess || occ < 1) { err = "linear-space kernel cannot be resident"; return 2; }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  size_t per = (size_t)a.lay.stride * sizeof(double);
  long long by_mem = (long long)(((double)(free_b + st->scratch_bytes) * 0.70) / (double)per);
  long long nslots = std::min<long long>(std::min<long long>(in.b.nseq, (long long)in.sm_count * occ), by_mem);
  if (in.max_slots > 0) nslots = std::min<long long>(nslots, in.max_slots);
  if (nslots < 1) { err = "not enough device memory for one sequence slot"; return 3; }
  size_t need = (size_t)nslots * per;
  if (need > st->scratch_bytes) {
    if (st->scratch) cudaFree(st->scratch);
    st->scratch = nullptr; st->scratch_bytes = 0;
    if (cudaMalloc(&st->scratch, need) != cudaSuccess) { err = "scratch allocation failed"; return 3; }
    st->scratch_bytes = need;
  }
  if (!st->queue && cudaMalloc(&st->queue, sizeof(int)) != cudaSuccess) { err = "queue allocation failed"; return 3; }
  cudaStream_t stream = (cudaStream_t)in.stream;
  cudaMemsetAsync(st->queue, 0, sizeof(int), stream);
  a.scratch = (double*)st->scratch; a.queue = (int*)st->queue;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  if (in.nch == 2) relem_estep_lin_kernel<2><<<(int)nslots, LIN_THREADS, a.lay.sm_total, stream>>>(a);
  else relem_estep_lin_kernel<1><<<(int)nslots, LIN_THREADS, a.lay.sm_total, stream>>>(a);
  e = cudaGetLastError();
  cudaEventRecord(e1, stream);
  if (e != cudaSuccess) { err = std::string("relem_estep_lin_kernel launch: ") + cudaGetErrorString(e); return 2; }
  e = cudaEventSynchronize(e1);
  if (e != cudaSuccess) { err = std::string("relem_estep_lin_kernel: ") + cudaGetErrorString(e); return 2; }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (kernel_ms) *kernel_ms = ms;
  if (launches) *launches = 1;
  return 0;
#endif
}

}  // namespace lin
}  // namespace relem
