// relem_lin.cu -- kernels and launcher of the scaled linear-space E-step (dp_lin.cuh).
//
// Execution model: WAVEFRONT ACROSS THE BATCH.  A chunk of sequences (as many as the scratch memory holds) is swept
// span by span; for every span d and every phase of the cell update one small kernel is launched whose grid covers
// (sequence, tile of cells on diagonal d) pairs, one warp per cell.  Kernel boundaries are the wavefront barriers.
// Why not one persistent CTA per sequence (the first version of this file): the complete cell update is ~12K SASS
// instructions, the SM's instruction cache holds ~2K, and ncu showed 2/3 of all stall samples as "no instruction"
// (profiles/r1_lin_persistent.md).  One phase is a few hundred instructions, every warp on the GPU runs the same
// phase at the same time, load balance is global instead of per CTA, and no intra-CTA barrier is left.
//
// Per chunk:  prep -> [energy-only inside d=3..W, exterior, exterior-outside, outside d=W..3, filter]
//             -> coupled inside d=0..W (phases L,P,B,E) -> exterior + Z -> exterior-outside
//             -> coupled outside d=W..0 (phases EM,B,P,L) -> fold counts.
// Per-sequence data (bases, weights, masks, accumulators) lives in a header in the sequence's scratch slot.
//
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

#ifdef RELEM_HOST_EMU
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
#endif

namespace relem {
namespace lin {
using namespace relem::dp;

#define LIN_THREADS 128
#define LIN_WARPS (LIN_THREADS / 32)
#define LIN_NPHASE_SLOTS 16

struct LinLayout {
  unsigned long long stride;  // doubles per slot
  unsigned long long aP, aE, aM, a1, a2, aLl, aLr, aO;
  unsigned long long bP, bEl, bEr, bM, bBl, bBr, b2, bL, bO, bch, boch;
  unsigned long long kP, kE, kM, k1, k2, kO, kbP, kbE, kbM, kbBl, kbBr, kb2, kbO, kPm, kbEm;
  // per-slot header
  unsigned long long hdr;    // [16]: Z^tt, Z^tf, Z^ft, bad, canonical pair count, L, ys, -, EH[nch*2]
  unsigned long long wsf;    // [Lmax+1] exp(position weight)
  unsigned long long cnt;    // [nch][ncnt] emission posterior sums per (list entry, base)
  unsigned long long masks;  // 4 x (Lmax+2)*mw words: bp, lf, bpr, lfr
  unsigned long long bytes;  // x [Lmax+2], sp3, sp4, sp6 [Lmax+1 each]
  unsigned long long post;   // scanner: linear start / inner / end posteriors, [Lmax+2] each
  unsigned long long expo;   // power-of-two exponents of the four exterior rows (K0 O, bO; coupled O, bO), [Lmax+2] each
  int Lmax, Wmax, mw, nch, ncnt, mask_words;
  int sm_ctx, sm_misc, sm_pcnt, sm_warp, warp_bytes_in, warp_bytes_inb, warp_bytes_out;
};

static LinLayout make_lin_layout(int Lmax, int max_span, const LinHMM& h, int nch) {
  LinLayout lay;
  std::memset(&lay, 0, sizeof(lay));
  int Wmax = Lmax < max_span ? Lmax : max_span;
  lay.Lmax = Lmax; lay.Wmax = Wmax; lay.mw = (Wmax + 1 + 31) / 32; lay.nch = nch;
  lay.ncnt = 5 * h.n_right + 5 * h.n_left + 25 * h.n_pair;
  lay.mask_words = (Lmax + 2) * lay.mw;
  unsigned long long cells = (unsigned long long)(Lmax + 1) * (Wmax + 1);
  unsigned long long band = cells * h.S, ext = (unsigned long long)(Lmax + 1) * h.S;
  unsigned long long o = 0;
  auto take = [&](unsigned long long n) { unsigned long long r = o; o += (n + 1) & ~1ull; return r; };
  lay.aP = take(band); lay.aE = take(band); lay.aM = take(band); lay.a1 = take(band); lay.a2 = take(band);
  lay.aLl = take(band); lay.aLr = take(band); lay.aO = take(ext);
  // scatter mode never reads E by its right end (only the right-flank gather did): no second E table
  lay.bP = take(band * nch); lay.bEl = take(band * nch); lay.bEr = take(LIN_SCATTER_ILOOP ? 0 : band * nch); lay.bM = take(band * nch);
  lay.bBl = take(band * nch); lay.bBr = take(band * nch); lay.b2 = take(band * nch); lay.bL = take(band * nch);
  lay.bO = take(ext * nch);
  lay.bch = band; lay.boch = ext;
  lay.kP = take(cells); lay.kE = take(cells); lay.kM = take(cells); lay.k1 = take(cells); lay.k2 = take(cells);
  lay.kO = take(Lmax + 1);
  lay.kbP = take(cells); lay.kbE = take(cells); lay.kbM = take(cells); lay.kbBl = take(cells); lay.kbBr = take(cells);
  lay.kb2 = take(cells); lay.kbO = take(Lmax + 1);
  lay.kPm = take(cells); lay.kbEm = take(cells);
  lay.hdr = take(16);
  lay.wsf = take(Lmax + 1);
  lay.cnt = take((unsigned long long)nch * lay.ncnt);
  lay.masks = take(2ull * lay.mask_words);
  lay.bytes = take((4ull * (Lmax + 2) + 7) / 8);
  lay.post = take(3ull * (Lmax + 2));
  lay.expo = take(4ull * (Lmax + 2));
  lay.stride = o;
  int b = 0;
  auto sm = [&](int bytes) { int r = b; b += (bytes + 15) & ~15; return r; };
  lay.sm_ctx = sm((int)sizeof(LinCtx));
  lay.sm_misc = sm(64);
  lay.sm_pcnt = b;
  lay.sm_warp = b;
  lay.warp_bytes_in = warp_lin_bytes(h.S, Wmax, 1, h.n_max, 0, 0, true);
  lay.warp_bytes_inb = warp_lin_bytes(h.S, Wmax, 1, h.n_max, 0, 0, true, true);   // inside B: + staging of the split gather
  lay.warp_bytes_out = warp_lin_bytes(h.S, Wmax, nch, lin_outside_nmax(h), h.n_right, h.n_left, false);
  return lay;
}

struct LinKArgs {
  BatchView b;
  int base, count;   // this chunk handles sequences order[base .. base+count), slot k <-> sequence base+k
  int d;             // diagonal of a phase kernel
  int tile, ntile;   // cells per CTA and CTAs per sequence of a phase kernel
  int win;           // 1: inside phase of the scanner's start-constrained pass -- tiles cover only the cells that
                     //    contain position Ys (all other cells keep the values of the unconstrained pass)
  LinLayout lay;
  double* scratch;
  const double* k0pow;
  int kp_n;          // entries of k0pow before the G table
  EstepOut out;
  ScanOut so;        // scanner runs only
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

#ifdef RELEM_HOST_EMU
#define LIN_KERNEL(T, MINB) inline void
#define LIN_SMEM_ARG , unsigned char* smem_raw, int emu_block
#define LIN_BLOCK_IDX emu_block
#define LIN_SHARED static
#else
#define LIN_KERNEL(T, MINB) __global__ void __launch_bounds__(T, MINB)
#define LIN_SMEM_ARG
#define LIN_BLOCK_IDX ((int)blockIdx.x)
#define LIN_SHARED __shared__
#endif

RDEV bool finite_pos(double v) { return v > 0. && v < (-NINF); }

// the CTA's view of sequence slot `sk`: LinCtx in shared memory pointing at the slot header
template <bool FROM_HDR = true>
RDEV LinCtx& lin_attach(const LinKArgs& a, unsigned char* smem_raw, int sk, double*& slot, int& n) {
  const LinLayout& lay = a.lay;
  LinCtx& c = *(LinCtx*)(smem_raw + lay.sm_ctx);
  slot = a.scratch + (unsigned long long)sk * lay.stride;
  n = a.b.order[a.base + sk];
  if (CTA_TID == 0) {
    const long long o = a.b.off[n];
    // the prep kernel leaves the length in the header: one dependent load less on the critical path of every CTA
    const int L = FROM_HDR ? (int)slot[lay.hdr + 5] : (int)(a.b.off[n + 1] - o);
    const int W = L < LC.en.max_span ? L : LC.en.max_span;
    const int C = W - 7 < LC.en.max_iloop ? W - 7 : LC.en.max_iloop;
    SeqView& q = c.q;
    q.L = L; q.W = W; q.C = C; q.W1 = W + 1; q.S = LC.h.S; q.cells = (unsigned)(L + 1) * (unsigned)(W + 1);
    q.mw = lay.mw; q.min_pair = 5; q.min_multi = 10; q.bpr = nullptr; q.lfr = nullptr;
    unsigned char* by = (unsigned char*)(slot + lay.bytes);
    q.x = by;
    q.sp3 = (signed char*)(by + (lay.Lmax + 2));
    q.sp4 = (signed char*)(by + (lay.Lmax + 2) + (lay.Lmax + 1));
    q.sp6 = (signed char*)(by + (lay.Lmax + 2) + 2 * (lay.Lmax + 1));
    unsigned* mk = (unsigned*)(slot + lay.masks);
    q.bp = mk; q.lf = mk + lay.mask_words;
    c.bpr = mk + 2 * lay.mask_words; c.lfr = mk + 3 * lay.mask_words;
    q.ws = a.b.ws + o; q.emit0 = nullptr; q.emitT = nullptr;
    c.wsf = slot + lay.wsf; c.k0pow = a.k0pow;
    c.Ceff = C < 30 ? C : 30;
    lin_outside_limits(W, LC.en.max_iloop, LC.en.no_ene != 0, c.Ceff, c.Csum, c.Cfl);
    c.ys = FROM_HDR ? (int)slot[lay.hdr + 6] : -1;
    c.pys = slot + lay.post; c.pyi = c.pys + (lay.Lmax + 2); c.pye = c.pyi + (lay.Lmax + 2);
  }
  CTA_SYNC();
  return c;
}

RDEV CTabs lin_tabs(const LinLayout& lay, double* slot) {
  CTabs t;
  t.aP = slot + lay.aP; t.aE = slot + lay.aE; t.aM = slot + lay.aM; t.a1 = slot + lay.a1; t.a2 = slot + lay.a2;
  t.aLl = slot + lay.aLl; t.aLr = slot + lay.aLr; t.aO = slot + lay.aO;
  t.bP = slot + lay.bP; t.bEl = slot + lay.bEl; t.bEr = slot + lay.bEr; t.bM = slot + lay.bM; t.bBl = slot + lay.bBl;
  t.bBr = slot + lay.bBr; t.b2 = slot + lay.b2; t.bL = slot + lay.bL; t.bO = slot + lay.bO;
  t.bch = (unsigned)lay.bch; t.boch = (unsigned)lay.boch;
  t.eO = slot + lay.expo + 2 * (lay.Lmax + 2); t.fO = t.eO + (lay.Lmax + 2);
  return t;
}
RDEV K0Tabs lin_k0tabs(const LinLayout& lay, double* slot, const double* G) {
  K0Tabs t0;
  t0.Pm = slot + lay.kPm; t0.bEm = slot + lay.kbEm; t0.G = G;
  t0.eO = slot + lay.expo; t0.fO = t0.eO + (lay.Lmax + 2);
  t0.P = slot + lay.kP; t0.E = slot + lay.kE; t0.M = slot + lay.kM; t0.o1 = slot + lay.k1; t0.o2 = slot + lay.k2;
  t0.O = slot + lay.kO; t0.bP = slot + lay.kbP; t0.bE = slot + lay.kbE; t0.bM = slot + lay.kbM;
  t0.bBl = slot + lay.kbBl; t0.bBr = slot + lay.kbBr; t0.b2 = slot + lay.kb2; t0.bO = slot + lay.kbO;
  return t0;
}

// separable part of the generic interior-loop weight (see K0Tabs::G)
LIN_KERNEL(LIN_THREADS, 8) relem_lin_gtab_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifdef RELEM_HOST_EMU
  (void)smem_raw; (void)emu_block;
#endif
  double* G = const_cast<double*>(a.k0pow) + a.kp_n;
  for (int t = CTA_TID; t < 1024; t += CTA_NTH) {
    int u1 = t >> 5, u2 = t & 31, du = u1 > u2 ? u1 - u2 : u2 - u1;
    double v = 0.;
    if (u1 >= 3 && u2 >= 3 && u1 + u2 <= 30)
      v = ld_ro(LC.el.internal + u1 + u2) * ld_ro(LC.el.ninio + du) * a.k0pow[u1 + u2];
    G[t] = v;
  }
}

// ------------------------------------------------------------------------------------------------ prep
// one CTA per sequence: bases, exp(position weights), special hairpins, canonical masks -> slot header
LIN_KERNEL(LIN_THREADS, 8) relem_lin_prep_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  LIN_SHARED int cnt_sh;
  const LinLayout& lay = a.lay;
  const int sk = LIN_BLOCK_IDX;
  if (sk >= a.count) return;
  double* slot; int n;
  LinCtx& c = lin_attach<false>(a, smem_raw, sk, slot, n);
  const SeqView& q = c.q;
  const int L = q.L;
  const long long o = a.b.off[n];
  unsigned char* x = (unsigned char*)(slot + lay.bytes);
  double* wsf = slot + lay.wsf;
  for (int t = CTA_TID; t < L; t += CTA_NTH) { x[t] = a.b.seq[o + t]; wsf[t] = exp(a.b.ws[o + t]); }
  if (CTA_TID == 0) { x[L] = 0; x[L + 1] = 0; }
  for (int t = CTA_TID; t < 16; t += CTA_NTH) slot[lay.hdr + t] = 0.;
  for (int t = CTA_TID; t < lay.nch * lay.ncnt; t += CTA_NTH) slot[lay.cnt + t] = 0.;
  for (int t = CTA_TID; t < 3 * (lay.Lmax + 2); t += CTA_NTH) slot[lay.post + t] = 0.;
  CTA_SYNC();
  cta_special_hairpins(LC.en, x, L, (signed char*)q.sp3, (signed char*)q.sp4, (signed char*)q.sp6);
  unsigned* mk = (unsigned*)(slot + lay.masks);
  unsigned* bp = mk; unsigned* lf = mk + lay.mask_words;
  cta_canonical_mask(q, bp);
  CTA_SYNC();
  cta_left_mask(q, bp, lf);
  const int total = cta_count_bits(bp, (L + 1) * lay.mw, &cnt_sh);
  cta_right_mask(q, bp, mk + 2 * lay.mask_words);
  CTA_SYNC();
  cta_right_mask(q, lf, mk + 3 * lay.mask_words);
  if (CTA_TID == 0) {
    slot[lay.hdr + 4] = (double)total;
    slot[lay.hdr + 5] = (double)L;
    slot[lay.hdr + 6] = -1.;
    if (a.out.bpp_eff) a.out.bpp_eff[n] = 1.;
  }
}

// ------------------------------------------------------------------------------------------------ phases
enum { PH_K0_IN = 0, PH_K0_OUT, PH_IN_L, PH_IN_P, PH_IN_B, PH_IN_E, PH_OUT_EM, PH_OUT_B, PH_OUT_P, PH_OUT_L,
       PH_OUT_LF, PH_OUT_LR,     // LF / LR: left / right flank gathers of outside L as kernels of their own
       PH_OUT_PQ,                // enclosing interior loops of outside P
       PH_OUT_ES };              // scatter mode: interior loops of E pushed down to P and the flanks (replaces PQ, LF, LR)

// posterior sums of the CTA -> the sequence's accumulators in its slot header
template <int NCH> RDEV void lin_flush_counts(const LinLayout& lay, double* slot, WarpLin& w, EhAcc<NCH>& eh) {
  const LinHMM& h = LC.h;
  double* g = slot + lay.cnt;
  w_sync();
  if (!LC.p.no_prf) {
    const int nR = 5 * h.n_right, nL = 5 * h.n_left;
    for (int ch = 0; ch < NCH; ++ch) {
      for (int r = lane_id(); r < nR; r += WARP_N) {
        double v = w.cntR[ch * nR + r];
        if (v != 0.) red_add(g + ch * lay.ncnt + r, v);
      }
      for (int r = lane_id(); r < nL; r += WARP_N) {
        double v = w.cntL[ch * nL + r];
        if (v != 0.) red_add(g + ch * lay.ncnt + nR + r, v);
      }
    }
  }
  for (int k = 0; k < NCH * 2; ++k) {
    double v = w_sum(eh.v[k]);
    if (lane_id() == 0 && v != 0.) red_add(slot + lay.hdr + 8 + k, v);
  }
}

// grid = count * ntile CTAs; CTA (sk, tk) owns cells [tk*tile, (tk+1)*tile) of diagonal d of sequence slot sk,
// its warps take them interleaved
// resident CTAs per SM the register allocation aims at: 8 for the energy-only phases (62 registers), 7 (72 registers)
// for the coupled ones, 8 (64 registers) for the split-gather phases, which are bound by memory latency
#ifndef LIN_MINB_SPLIT
#define LIN_MINB_SPLIT 8   // measured: 9 314 -> 9 446 sequence-evaluations/s (64 registers, no spills)
#endif
#ifndef LIN_MINB_SMALL
#define LIN_MINB_SMALL 7
#endif
#ifndef LIN_FUSE_CELLS
// chunk size (sequences x longest length) up to which the phases of a diagonal are fused into one launch.  Measured
// (tools/minibatch_probe.py, 200-nt reads, two lanes): 64 sequence-evaluations 18.4 -> 16.3 ms fused, 128: 25.4 -> 26.7,
// 512: 68 -> 99 (a warp that owns a heavy cell now carries all of its phases), so only chunks of <= ~40 reads fuse.
#define LIN_FUSE_CELLS 8192
#endif
#ifndef LIN_STRIDED_SPARSE
#define LIN_STRIDED_SPARSE 1   // A/B switch: cells of the sparse phases interleaved over the CTAs of a sequence
#endif
template <int PH, int NCH, int MODE = 0>
LIN_KERNEL(LIN_THREADS, (PH <= PH_K0_OUT ? 8 : (PH == PH_OUT_B || PH == PH_IN_B) ? LIN_MINB_SPLIT : (PH == PH_OUT_L || PH == PH_OUT_EM || PH == PH_IN_P || PH == PH_IN_L) ? LIN_MINB_SMALL : 7))
relem_lin_phase_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const LinLayout& lay = a.lay;
  const int blk = LIN_BLOCK_IDX;
  const int sk = blk / a.ntile, tk = blk - sk * a.ntile;
  if (sk >= a.count) return;
  double* slot = a.scratch + (unsigned long long)sk * lay.stride;
  if (slot[lay.hdr + 3] != 0.) return;
  int n;
  LinCtx& c = lin_attach(a, smem_raw, sk, slot, n);
  const SeqView& q = c.q;
  const int d = a.d;
  if (d > q.W) return;
  const int ncell = q.L + 1 - d;
  int i0 = tk * a.tile, i1 = ncell < i0 + a.tile ? ncell : i0 + a.tile;
  // Scanner, second (start-constrained) pass.  The constraint vetoes emissions AT position Ys only, so
  //  * inside: a cell that does not contain Ys has bit for bit the value the unconstrained pass left in the table;
  //    only cells with i <= Ys < i+d are recomputed (one position of margin on either side);
  //  * outside (MODE 2 collects end posteriors only): a cell whose bases all lie before Ys can neither emit a motif
  //    end nor be read by a cell that can -- readers of an outside value are sub-cells -- so it is skipped.
  if (PH >= PH_IN_L && PH <= PH_IN_E && a.win) {
    const int ys = c.ys;
    const int lo = ys - d > 0 ? ys - d : 0, hi = ys + 1 < ncell - 1 ? ys + 1 : ncell - 1;
    i0 = lo + tk * a.tile;
    i1 = hi + 1 < i0 + a.tile ? hi + 1 : i0 + a.tile;
  }
  if (PH >= PH_OUT_EM && MODE == 2) {
    const int lo = c.ys - 1 - d;
    if (i0 < lo) i0 = lo;
  }
  const int w0 = warp_id(), nw = n_warps();
  // Phases that touch few, unevenly spread cells (pairs and what they enclose): instead of a contiguous tile the CTA
  // takes every ntile-th group of cells over the whole range, so stems do not pile up in one CTA (ncu: 19 % warps active
  // in the interior-loop scatter with contiguous tiles, profiles/r2_phase_kernels.md).
  constexpr bool kStrided = LIN_STRIDED_SPARSE && (PH == PH_IN_P || PH == PH_IN_E || PH == PH_OUT_EM || PH == PH_OUT_P ||
                                                   PH == PH_OUT_ES || PH == PH_OUT_PQ);
  int first = i0 + w0, step = nw;
  if (kStrided) {
    int r0 = 0, r1 = ncell;
    if (PH >= PH_IN_L && PH <= PH_IN_E && a.win) {
      r0 = c.ys - d > 0 ? c.ys - d : 0;
      r1 = (c.ys + 1 < ncell - 1 ? c.ys + 1 : ncell - 1) + 1;
    }
    if (PH >= PH_OUT_EM && MODE == 2 && r0 < c.ys - 1 - d) r0 = c.ys - 1 - d;
    first = r0 + tk * nw + w0; step = a.ntile * nw; i1 = r1;
    if (r0 + tk * nw >= r1) return;
  } else if (i0 >= i1) return;
  if (PH == PH_K0_IN || PH == PH_K0_OUT) {
    K0Tabs t0 = lin_k0tabs(lay, slot, a.k0pow + a.kp_n);
    for (int i = i0 + w0; i < i1; i += nw) {
      int* sbuf = (int*)(smem_raw + lay.sm_warp) + w0 * 128;
      if (PH == PH_K0_IN) k0_inside_cell(c, t0, i, d, sbuf);
      else k0_outside_cell(c, t0, i, d, sbuf);
    }
    return;
  }
  CTabs t = lin_tabs(lay, slot);
  const LinHMM& h = LC.h;
  if (PH >= PH_IN_L && PH <= PH_IN_E) {
    WarpLin w = warp_lin_carve(smem_raw + lay.sm_warp + w0 * (PH == PH_IN_B ? lay.warp_bytes_inb : lay.warp_bytes_in), q.S, lay.Wmax,
                               1, h.n_max, 0, 0, true, PH == PH_IN_B);
#if LIN_SPLIT_TMA && !defined(RELEM_HOST_EMU)
    if (PH == PH_IN_B) {   // the warp's bulk-copy barrier: one arrival (the lane that arms it) per batch
      if (lane_id() == 0) {
        mbar_init(w.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      w_sync();
    }
#endif
    for (int i = first; i < i1; i += step) {
      if (PH == PH_IN_L) lin_in_L(c, t, i, d, w);
      if (PH == PH_IN_P) { if (ok_P(q, i, d)) lin_in_P(c, t, i, d, w); }
      if (PH == PH_IN_B) {
        bool gB = ok_B(q, i, d), gM = ok_M(q, i, d);
        if (gB || gM) lin_in_B(c, t, i, d, ok_P(q, i, d), gB, gM, w);
      }
      if (PH == PH_IN_E) { if (ok_E(q, i, d)) lin_in_E(c, t, i, d, ok_M(q, i, d), w); }
    }
    return;
  }
  if (PH >= PH_OUT_EM) {
    WarpLin w = warp_lin_carve(smem_raw + lay.sm_warp + w0 * lay.warp_bytes_out, q.S, lay.Wmax, NCH, lin_outside_nmax(h), h.n_right,
                               h.n_left, false);
    w.pcnt = slot + lay.cnt + 5 * (h.n_right + h.n_left);
    w.pstride = (unsigned)lay.ncnt;
    for (int tt = lane_id(); tt < NCH * 5 * h.n_right; tt += WARP_N) w.cntR[tt] = 0.;
    for (int tt = lane_id(); tt < NCH * 5 * h.n_left; tt += WARP_N) w.cntL[tt] = 0.;
    w_sync();
    EhAcc<NCH> eh;
    for (int k = 0; k < NCH * 2; ++k) eh.v[k] = 0.;
    for (int i = first; i < i1; i += step) {
      if (PH == PH_OUT_EM) {
        bool gE = ok_E(q, i, d), gM = ok_M(q, i, d);
        if (gE || gM) lin_out_EM<NCH, MODE>(c, t, i, d, gE, gM, w, eh);
      }
      if (PH == PH_OUT_B) { if (ok_B(q, i, d)) lin_out_B<NCH, MODE>(c, t, i, d, ok_M(q, i, d), w); }
      if (PH == PH_OUT_P) { if (ok_P(q, i, d)) lin_out_P<NCH, MODE, 1>(c, t, i, d, ok_B(q, i, d), w, eh); }
      if (PH == PH_OUT_PQ) { if (ok_P(q, i, d)) lin_out_P<NCH, MODE, 2>(c, t, i, d, false, w, eh); }
      if (PH == PH_OUT_ES) { if (ok_E(q, i, d)) lin_out_ES<NCH, MODE>(c, t, i, d, w, eh); }
      if (PH == PH_OUT_L) lin_out_L<NCH, MODE, 1>(c, t, i, d, d >= 3 && ok_E(q, i, d), w, eh);
      if (PH == PH_OUT_LF) lin_out_L<NCH, MODE, 2>(c, t, i, d, false, w, eh);
      if (PH == PH_OUT_LR) lin_out_L<NCH, MODE, 3>(c, t, i, d, false, w, eh);
    }
    if (PH != PH_OUT_LF && PH != PH_OUT_LR) lin_flush_counts<NCH>(lay, slot, w, eh);
  }
}

// ------------------------------------------------------------------------------------------------ fused diagonal
// A training minibatch (~128 sequences) cannot fill 1036 resident CTAs phase by phase: each of the ~9 launches per
// diagonal is a fraction of a wave and ends on its slowest cell.  For small chunks one launch per diagonal and direction
// runs every phase of a cell in the warp that owns it (the phases of one cell depend only on each other and on other
// diagonals), cells interleaved over the CTAs.  DIR 0 = inside (L, P, B, E), DIR 1 = outside (E/M, interior-loop
// scatter, B, P, L).  Trainer only (no scan windows).
template <int DIR, int NCH>
LIN_KERNEL(LIN_THREADS, 7)
relem_lin_diag_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const LinLayout& lay = a.lay;
  const int blk = LIN_BLOCK_IDX;
  const int sk = blk / a.ntile, tk = blk - sk * a.ntile;
  if (sk >= a.count) return;
  double* slot = a.scratch + (unsigned long long)sk * lay.stride;
  if (slot[lay.hdr + 3] != 0.) return;
  int n;
  LinCtx& c = lin_attach(a, smem_raw, sk, slot, n);
  const SeqView& q = c.q;
  const int d = a.d;
  if (d > q.W) return;
  const int ncell = q.L + 1 - d;
  const int w0 = warp_id(), nw = n_warps();
  if (tk * nw >= ncell) return;
  const int first = tk * nw + w0, step = a.ntile * nw;
  CTabs t = lin_tabs(lay, slot);
  const LinHMM& h = LC.h;
  if (DIR == 0) {
    const int wb = lay.warp_bytes_in > lay.warp_bytes_inb ? lay.warp_bytes_in : lay.warp_bytes_inb;
    WarpLin w = warp_lin_carve(smem_raw + lay.sm_warp + w0 * wb, q.S, lay.Wmax, 1, h.n_max, 0, 0, true, false);
    for (int i = first; i < ncell; i += step) {
      lin_in_L(c, t, i, d, w);
      if (d >= 5) {
        const bool gP = ok_P(q, i, d), gB = ok_B(q, i, d), gM = ok_M(q, i, d);
        if (gP) lin_in_P(c, t, i, d, w);
        if (gB || gM) lin_in_B(c, t, i, d, gP, gB, gM, w);
      }
      if (d >= 3 && ok_E(q, i, d)) lin_in_E(c, t, i, d, ok_M(q, i, d), w);
    }
    return;
  }
  WarpLin w = warp_lin_carve(smem_raw + lay.sm_warp + w0 * lay.warp_bytes_out, q.S, lay.Wmax, NCH, lin_outside_nmax(h), h.n_right,
                             h.n_left, false);
  w.pcnt = slot + lay.cnt + 5 * (h.n_right + h.n_left);
  w.pstride = (unsigned)lay.ncnt;
  for (int tt = lane_id(); tt < NCH * 5 * h.n_right; tt += WARP_N) w.cntR[tt] = 0.;
  for (int tt = lane_id(); tt < NCH * 5 * h.n_left; tt += WARP_N) w.cntL[tt] = 0.;
  w_sync();
  EhAcc<NCH> eh;
  for (int k = 0; k < NCH * 2; ++k) eh.v[k] = 0.;
  for (int i = first; i < ncell; i += step) {
    const bool gE = d >= 3 && ok_E(q, i, d), gM = d >= 3 && ok_M(q, i, d);
    if (gE || gM) lin_out_EM<NCH, 0>(c, t, i, d, gE, gM, w, eh);
    if (gE) lin_out_ES<NCH, 0>(c, t, i, d, w, eh);
    if (d >= 5) {
      const bool gB = ok_B(q, i, d);
      if (gB) lin_out_B<NCH, 0>(c, t, i, d, gM, w);
      if (ok_P(q, i, d)) lin_out_P<NCH, 0, 1>(c, t, i, d, gB, w, eh);
    }
    lin_out_L<NCH, 0, 1>(c, t, i, d, gE, w, eh);
  }
  lin_flush_counts<NCH>(lay, slot, w, eh);
}

// ------------------------------------------------------------------------------------------------ zero
// scatter mode: b^P (every cell) and b^L (spans 1..Ceff, the possible flanks) start the outside pass at zero, because
// lin_out_ES adds to them before their own phase runs.  128-bit stores over the contiguous b^P plane.
template <int NCH> LIN_KERNEL(LIN_THREADS, 8) relem_lin_zero_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifdef RELEM_HOST_EMU
  (void)smem_raw;
#endif
  const LinLayout& lay = a.lay;
  const int blk = LIN_BLOCK_IDX;
  const int sk = blk / a.ntile, tk = blk - sk * a.ntile;
  if (sk >= a.count) return;
  double* slot = a.scratch + (unsigned long long)sk * lay.stride;
  if (slot[lay.hdr + 3] != 0.) return;
  // the band tables of a slot are laid out with the sequence's OWN span W = min(L, max-span) (cidx), not the chunk's
  const int L = (int)slot[lay.hdr + 5];
  const int W = L < LC.en.max_span ? L : LC.en.max_span;
  const int S = LC.h.S, W1 = W + 1;
  const int C = W - 7 < LC.en.max_iloop ? W - 7 : LC.en.max_iloop;
  int Ceff = C < 30 ? C : 30, csum_unused;
  lin_outside_limits(W, LC.en.max_iloop, LC.en.no_ene != 0, Ceff, csum_unused, Ceff);   // flanks the scatter can reach
  // b^P: bch doubles per channel, even and 16-byte aligned (slot offsets are even, the scratch is 256-byte aligned)
  const unsigned long long nP = (unsigned long long)NCH * lay.bch / 2;
  double2* pP = reinterpret_cast<double2*>(slot + lay.bP);
  for (unsigned long long z = (unsigned long long)tk * CTA_NTH + CTA_TID; z < nP; z += (unsigned long long)a.ntile * CTA_NTH)
    pP[z] = make_double2(0., 0.);
  if ((lay.bch & 1ull) && tk == 0 && CTA_TID == 0)
    for (int ch = 0; ch < NCH; ++ch) slot[lay.bP + (unsigned long long)(ch + 1) * lay.bch - 1] = 0.;
  if (Ceff >= 1) {
    const unsigned per_row = (unsigned)Ceff * S;
    const unsigned long long nL = (unsigned long long)(L + 1) * per_row;
    for (int ch = 0; ch < NCH; ++ch) {
      double* pL = slot + lay.bL + (unsigned long long)ch * lay.bch;
      for (unsigned long long z = (unsigned long long)tk * CTA_NTH + CTA_TID; z < nL; z += (unsigned long long)a.ntile * CTA_NTH) {
        unsigned row = (unsigned)(z / per_row), rem = (unsigned)(z - (unsigned long long)row * per_row);
        pL[((unsigned long long)row * W1 + 1) * S + rem] = 0.;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ exterior rows
// one warp per sequence.  WHICH: 0 energy-only inside, 1 energy-only outside, 2 coupled inside (+ partition functions,
// root weights), 3 coupled outside
template <int WHICH, int NCH, int MODE = 0> LIN_KERNEL(32, 16) relem_lin_ext_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const LinLayout& lay = a.lay;
  const int sk = LIN_BLOCK_IDX;
  if (sk >= a.count) return;
  double* slot = a.scratch + (unsigned long long)sk * lay.stride;
  if (slot[lay.hdr + 3] != 0.) return;
  int n;
  LinCtx& c = lin_attach(a, smem_raw, sk, slot, n);
  const SeqView& q = c.q;
  const LinHMM& h = LC.h;
  const int L = q.L, S = q.S;
  if (WHICH == 0) {
    K0Tabs t0 = lin_k0tabs(lay, slot, a.k0pow + a.kp_n);
    k0_inside_ext(c, t0);
    w_sync();
    const double Z0 = ld_cg(t0.O + L);
    if (!finite_pos(Z0)) {
      if (lane_id() == 0) { slot[lay.hdr + 3] = 1.; a.flag[n] = 1; }
    }
    return;
  }
  if (WHICH == 1) {
    K0Tabs t0 = lin_k0tabs(lay, slot, a.k0pow + a.kp_n);
    k0_outside_ext(c, t0, 1. / ld_cg(t0.O + L));
    return;
  }
  CTabs t = lin_tabs(lay, slot);
  if (WHICH == 2) {
    WarpLin w = warp_lin_carve(smem_raw + lay.sm_warp, S, lay.Wmax, 1, h.n_max, 0, 0, true);
    lin_inside_ext(c, t, w);
    w_sync();
    const double r00 = h.s00 >= 0 ? ld_cg(t.aO + (unsigned)L * S + h.s00) : 0.;
    const double rM2 = h.s0M2 >= 0 ? ld_cg(t.aO + (unsigned)L * S + h.s0M2) : 0.;
    const double rM1 = h.s0M1 >= 0 ? ld_cg(t.aO + (unsigned)L * S + h.s0M1) : 0.;
    const double Ztt = r00 + (rM2 + rM1), Ztf = rM2 + rM1, Zft = r00;
    double rw[NCH][3];
    if (MODE == 0) {
      const int kind = a.b.kind[n];
      // every partition function the trainer tests must be representable; otherwise the log-space path decides
      const bool bad = !finite_pos(Ztt) || (kind != 2 && !finite_pos(Ztf)) || !(Zft >= 0. && Zft < (-NINF));
      if (bad) {
        if (lane_id() == 0) { slot[lay.hdr + 3] = 1.; a.flag[n] = 1; }
        return;
      }
      if (lane_id() == 0) {
        slot[lay.hdr + 0] = Ztt; slot[lay.hdr + 1] = Ztf; slot[lay.hdr + 2] = Zft;
        // ln Z = ln(mantissa) + eO(L) ln 2 - L ln kappa
        const double shift = -(double)L * LC.p.ln_kappa + t.eO[L] * 0.6931471805599453;
        a.out.Z[n * 3 + 0] = log(Ztt) + shift;
        a.out.Z[n * 3 + 1] = Ztf > 0. ? log(Ztf) + shift : NINF;
        a.out.Z[n * 3 + 2] = Zft > 0. ? log(Zft) + shift : NINF;
        a.out.skipped[n] = 0;
      }
      // root weights of the outside pass: channel 0 = Zo (all three roots), channel 1 = the restricted condition;
      // NCH = 1 carries their difference
      double o0 = 1. / Ztt;
      double x00 = 0., xM = 0.;
      if (kind == 1 || kind >= 3) xM = 1. / Ztf;
      else x00 = Zft > 0. ? 1. / Zft : 0.;
      if (NCH == 2) {
        rw[0][0] = o0; rw[0][1] = o0; rw[0][2] = o0;
        rw[NCH - 1][0] = x00; rw[NCH - 1][1] = xM; rw[NCH - 1][2] = xM;
      } else {
        rw[0][0] = o0 - x00; rw[0][1] = o0 - xM; rw[0][2] = o0 - xM;
      }
    } else if (MODE == 1) {
      // scanner, unconstrained pass: posteriors relative to Z(1,1) (calc_motif_start_position, motif_scanner.hpp:186-193)
      if (!finite_pos(Ztt)) {
        if (lane_id() == 0) { slot[lay.hdr + 3] = 1.; a.flag[n] = 1; }
        return;
      }
      if (lane_id() == 0 && a.so.ZL) a.so.ZL[n] = log(Ztt) - (double)L * LC.p.ln_kappa + t.eO[L] * 0.6931471805599453;
      for (int k = 0; k < 3; ++k) rw[0][k] = 1. / Ztt;
    } else {
      // scanner, start fixed: a sequence in which the motif cannot start anywhere has Z = 0 and no end posterior
      const double o0 = finite_pos(Ztt) ? 1. / Ztt : 0.;
      for (int k = 0; k < 3; ++k) rw[0][k] = o0;
    }
    for (int tt = lane_id(); tt < NCH * S; tt += WARP_N) {
      int ch = tt / S, s = tt - ch * S;
      double v = 0.;
      if (s == h.s00) v = rw[ch][0];
      if (s == h.s0M2) v = rw[ch][1];
      if (s == h.s0M1) v = rw[ch][2];
      t.bO[ch * t.boch + (unsigned)L * S + s] = v;
    }
    if (lane_id() == 0) t.fO[L] = -t.eO[L];   // root weights are 1 / mantissa sums
    return;
  }
  if (WHICH == 3) {
    WarpLin w = warp_lin_carve(smem_raw + lay.sm_warp, S, lay.Wmax, NCH, lin_outside_nmax(h), h.n_right, h.n_left, false);
    w.pcnt = slot + lay.cnt + 5 * (h.n_right + h.n_left);
    w.pstride = (unsigned)lay.ncnt;
    for (int tt = lane_id(); tt < NCH * 5 * h.n_right; tt += WARP_N) w.cntR[tt] = 0.;
    for (int tt = lane_id(); tt < NCH * 5 * h.n_left; tt += WARP_N) w.cntL[tt] = 0.;
    w_sync();
    EhAcc<NCH> eh;
    for (int k = 0; k < NCH * 2; ++k) eh.v[k] = 0.;
    lin_outside_ext<NCH, MODE>(c, t, w);
    lin_flush_counts<NCH>(lay, slot, w, eh);
  }
}

// ------------------------------------------------------------------------------------------------ filter
// one CTA per sequence: keep pairs with ln BPP >= ln min_bpp (energy_model.hpp:257-261), rebuild the four masks
LIN_KERNEL(LIN_THREADS, 8) relem_lin_filter_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  LIN_SHARED int cnt_sh;
  const LinLayout& lay = a.lay;
  const int sk = LIN_BLOCK_IDX;
  if (sk >= a.count) return;
  double* slot = a.scratch + (unsigned long long)sk * lay.stride;
  if (slot[lay.hdr + 3] != 0.) return;
  int n;
  LinCtx& c = lin_attach(a, smem_raw, sk, slot, n);
  const SeqView& q = c.q;
  const int L = q.L;
  K0Tabs t0 = lin_k0tabs(lay, slot, a.k0pow + a.kp_n);
  unsigned* mk = (unsigned*)(slot + lay.masks);
  unsigned* bp = mk; unsigned* lf = mk + lay.mask_words;
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
  const int nbp = cta_count_bits(bp, (L + 1) * lay.mw, &cnt_sh);
  cta_right_mask(q, bp, mk + 2 * lay.mask_words);
  CTA_SYNC();
  cta_right_mask(q, lf, mk + 3 * lay.mask_words);
  if (CTA_TID == 0 && a.out.bpp_eff) a.out.bpp_eff[n] = (double)nbp / slot[lay.hdr + 4];
}

// ------------------------------------------------------------------------------------------------ fold
// one CTA per sequence: per-entry emission sums -> theta-shaped counts; results
template <int NCH> LIN_KERNEL(LIN_THREADS, 8) relem_lin_fold_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const LinLayout& lay = a.lay;
  const LinHMM& h = LC.h;
  const int NT = LC.p.n_theta;
  const int sk = LIN_BLOCK_IDX;
  if (sk >= a.count) return;
  double* slot = a.scratch + (unsigned long long)sk * lay.stride;
  if (slot[lay.hdr + 3] != 0.) return;
  const int n = a.b.order[a.base + sk];
  double* sen = (double*)smem_raw;  // [NCH][NT]
  for (int tt = CTA_TID; tt < NCH * NT; tt += CTA_NTH) sen[tt] = 0.;
  CTA_SYNC();
  const double* g = slot + lay.cnt;
  const int nR = 5 * h.n_right, nL = 5 * h.n_left;
  for (int tt = CTA_TID; tt < NCH * lay.ncnt; tt += CTA_NTH) {
    int ch = tt / lay.ncnt, r = tt - ch * lay.ncnt;
    double v = ld_cg(g + tt);
    if (v == 0.) continue;
    if (r < nR) { int idx = ld_ro(h.r_en + r); if (idx >= 0) sm_add(sen + ch * NT + idx, v); }
    else if (r < nR + nL) { int idx = ld_ro(h.l_en + (r - nR)); if (idx >= 0) sm_add(sen + ch * NT + idx, v); }
    else {
      int i1 = ld_ro(h.p_en1 + (r - nR - nL)), i2 = ld_ro(h.p_en2 + (r - nR - nL));
      if (i1 >= 0) sm_add(sen + ch * NT + i1, v);
      if (i2 >= 0) sm_add(sen + ch * NT + i2, v);
    }
  }
  CTA_SYNC();
  // non-finite counts mean the scaled tables overflowed somewhere: let the log-space path redo the sequence
  bool okv = true;
  for (int tt = 0; tt < NCH * NT; ++tt) okv = okv && (sen[tt] - sen[tt] == 0.);
  double ehv[NCH * 2];
  for (int k = 0; k < NCH * 2; ++k) { ehv[k] = ld_cg(slot + lay.hdr + 8 + k); okv = okv && (ehv[k] - ehv[k] == 0.); }
  if (!okv) {
    if (CTA_TID == 0) a.flag[n] = 1;
    return;
  }
  for (int tt = CTA_TID; tt < NT; tt += CTA_NTH) {
    a.out.ENo[(long long)n * NT + tt] = sen[tt];
    a.out.ENx[(long long)n * NT + tt] = NCH == 2 ? sen[(NCH - 1) * NT + tt] : 0.;
  }
  if (CTA_TID == 0) {
    a.out.EH[n * 4 + 0] = ehv[0]; a.out.EH[n * 4 + 1] = ehv[1];
    a.out.EH[n * 4 + 2] = NCH == 2 ? ehv[(NCH - 1) * 2] : 0.;
    a.out.EH[n * 4 + 3] = NCH == 2 ? ehv[(NCH - 1) * 2 + 1] : 0.;
  }
}

// ------------------------------------------------------------------------------------------------ scanner folds
// index of the last maximum (max_index, util.hpp:231-241); NaN never wins
RDEV int lin_last_max(const double* v, int n) {
  int s = 0;
  double mx = -1.7976931348623157e308;
  for (int i = 0; i < n; ++i)
    if (mx <= v[i]) { s = i; mx = v[i]; }
  return s;
}
// STAGE 1 (after the unconstrained pass): E[N], log start / inner posteriors, exist prob, Ys (-> header).
// STAGE 2 (after the start-constrained pass): log end posteriors, Ye.   One CTA per sequence.
template <int STAGE> LIN_KERNEL(LIN_THREADS, 8) relem_lin_scanfold_kernel(LinKArgs a LIN_SMEM_ARG) {
#ifndef RELEM_HOST_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  const LinLayout& lay = a.lay;
  const LinHMM& h = LC.h;
  const int NT = LC.p.n_theta;
  const int sk = LIN_BLOCK_IDX;
  if (sk >= a.count) return;
  double* slot = a.scratch + (unsigned long long)sk * lay.stride;
  if (slot[lay.hdr + 3] != 0.) return;
  const int n = a.b.order[a.base + sk];
  const long long o = a.b.off[n];
  const int L = (int)slot[lay.hdr + 5];
  double* pys = slot + lay.post; double* pyi = pys + (lay.Lmax + 2); double* pye = pyi + (lay.Lmax + 2);
  if (STAGE == 1) {
    double* sen = (double*)smem_raw;  // [NT]
    for (int tt = CTA_TID; tt < NT; tt += CTA_NTH) sen[tt] = 0.;
    CTA_SYNC();
    const double* g = slot + lay.cnt;
    const int nR = 5 * h.n_right, nL = 5 * h.n_left;
    for (int r = CTA_TID; r < lay.ncnt; r += CTA_NTH) {
      double v = ld_cg(g + r);
      if (v == 0.) continue;
      if (r < nR) { int idx = ld_ro(h.r_en + r); if (idx >= 0) sm_add(sen + idx, v); }
      else if (r < nR + nL) { int idx = ld_ro(h.l_en + (r - nR)); if (idx >= 0) sm_add(sen + idx, v); }
      else {
        int i1 = ld_ro(h.p_en1 + (r - nR - nL)), i2 = ld_ro(h.p_en2 + (r - nR - nL));
        if (i1 >= 0) sm_add(sen + i1, v);
        if (i2 >= 0) sm_add(sen + i2, v);
      }
    }
    CTA_SYNC();
    bool okv = true;
    for (int tt = 0; tt < NT; ++tt) okv = okv && (sen[tt] - sen[tt] == 0.);
    if (!okv) {
      if (CTA_TID == 0) { slot[lay.hdr + 3] = 1.; a.flag[n] = 1; }
      return;
    }
    for (int tt = CTA_TID; tt < NT; tt += CTA_NTH) a.so.EN[(long long)n * NT + tt] = sen[tt];
    for (int tt = CTA_TID; tt < L; tt += CTA_NTH) {
      double s0 = ld_cg(pys + tt), s1 = ld_cg(pyi + tt);
      double ls = s0 > 0. ? log(s0) : NINF;
      pys[tt] = ls;
      a.so.PysL[o + tt] = ls;
      a.so.PyiL[o + tt] = s1 > 0. ? log(s1) : NINF;
    }
    CTA_SYNC();
    if (CTA_TID == 0) {
      const int ys = lin_last_max(pys, L);
      double ex = 0.;  // exp(sumL(PysL)), motif_scanner.hpp:246
      for (int tt = 0; tt < L; ++tt) if (pys[tt] > NINF) ex += exp(pys[tt]);
      a.so.exist[n] = ex;
      a.so.Ys[n] = ys;
      slot[lay.hdr + 6] = (double)ys;
    }
  } else {
    for (int tt = CTA_TID; tt <= L; tt += CTA_NTH) {
      double e0 = ld_cg(pye + tt);
      double le = e0 > 0. ? log(e0) : NINF;
      pye[tt] = le;
      a.so.PyeL[o + n + tt] = le;
    }
    CTA_SYNC();
    if (CTA_TID == 0) a.so.Ye[n] = lin_last_max(pye, L + 1);
  }
}

// ------------------------------------------------------------------------------------------------ launcher
struct LinState {
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* k0pow = nullptr;
  size_t k0pow_n = 0;
  double k0pow_kappa = 0.;   // what the device copy of the kappa0 powers was built for
  int k0pow_kp = 0;
#ifndef RELEM_HOST_EMU
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t lane[4] = {nullptr, nullptr, nullptr, nullptr};  // two chunks in flight: one fills the SMs while the other's kernel drains
  cudaEvent_t lane_done[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> ev_pool;   // RELEM_PHASE_TIMING: one event pair per phase launch
#endif
  float phase_ms[LIN_NPHASE_SLOTS] = {0};   // device time per phase class of the last instrumented launch
  int phase_launches[LIN_NPHASE_SLOTS] = {0};
};

LinState* lin_state_create() { return new LinState(); }
void lin_state_destroy(LinState* s) {
  if (!s) return;
#ifdef RELEM_HOST_EMU
  std::free(s->scratch); std::free(s->k0pow);
#else
  if (s->scratch) cudaFree(s->scratch);
  if (s->k0pow) cudaFree(s->k0pow);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  for (int k = 0; k < 4; ++k) {
    if (s->lane[k]) cudaStreamDestroy(s->lane[k]);
    if (s->lane_done[k]) cudaEventDestroy(s->lane_done[k]);
  }
  for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
#endif
  delete s;
}

namespace {
struct Runner {
  LinKArgs a;
  int smem_in, smem_inb, smem_out, smem_small, smem_k0, smem_ext_in, smem_ext_out;
  int launches = 0;
  int resident_ctas = 148 * 7;
  int cmax = 30;                   // longest unpaired flank of an interior loop: min(30, max_iloop, W - 7)
  int fill = 4;                    // CTAs per resident slot a small launch aims for (RELEM_FILL)
  bool fuse = false;               // chunk small enough for one launch per diagonal (RELEM_FUSE_SEQS)
  int tile_p = 128, tile_e = 256;   // cells per CTA of the phases that touch few cells
  int tile_k0 = 32, tile_d = 64;   // cells per CTA of the dense phases (measured sweet spot, RELEM_TILE_*)
  int tile_f = 64, tile_q = 256;   // flank kernels of outside L, interior-loop kernel of outside P
#ifdef RELEM_HOST_EMU
  std::vector<unsigned char> smem;
  void mark(int) {}
#else
  cudaStream_t stream;
  // RELEM_PHASE_TIMING=1 (one lane): an event before and after every phase launch, summed per phase class afterwards
  LinState* st = nullptr;
  bool timing = false;
  std::vector<int> marks;   // phase id of event pair k (events 2k, 2k+1 of the pool)
  void mark(int ph) {
    if (!timing) return;
    const size_t k = marks.size() * 2 + (open_ ? 1 : 0);
    while (st->ev_pool.size() <= k) { cudaEvent_t e; cudaEventCreate(&e); st->ev_pool.push_back(e); }
    cudaEventRecord(st->ev_pool[k], stream);
    if (open_) marks.push_back(ph);
    open_ = !open_;
  }
  bool open_ = false;
#endif
};
}  // namespace

#ifdef RELEM_HOST_EMU
#define LIN_LAUNCH(R, KERN, GRID, THREADS, SMEM) \
  do { for (int b__ = 0; b__ < (GRID); ++b__) KERN((R).a, (R).smem.data(), b__); ++(R).launches; } while (0)
#else
#define LIN_LAUNCH(R, KERN, GRID, THREADS, SMEM)                                                              \
  do {                                                                                                        \
    if ((SMEM) > 48 * 1024)                                                                                   \
      cudaFuncSetAttribute((const void*)KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (SMEM));           \
    KERN<<<(GRID), (THREADS), (SMEM), (R).stream>>>((R).a);                                                   \
    ++(R).launches;                                                                                           \
  } while (0)
#endif

// Small chunks (a training minibatch is ~128 sequences) cannot fill the GPU with 64-cell tiles: shrink the tile until
// the launch has ~2 CTAs per resident slot, down to one cell per warp.
static int fit_tile(const Runner& r, int tile, int ncell_max) {
  const long long want = (long long)r.fill * r.resident_ctas;
  long long t = ((long long)r.a.count * ncell_max + want - 1) / want;   // cells per CTA that give `want` CTAs
  t = (t + LIN_WARPS - 1) / LIN_WARPS * LIN_WARPS;
  if (t < LIN_WARPS) t = LIN_WARPS;
  return (int)std::min<long long>(tile, t);
}

template <int PH, int NCH> static void launch_phase(Runner& r, int d, int tile, int smem) {
  const int ncell_max = r.a.lay.Lmax + 1 - d;
  if (ncell_max <= 0) return;
  if (tile > ncell_max) tile = ncell_max;
  tile = fit_tile(r, tile, ncell_max);
  r.a.d = d; r.a.tile = tile; r.a.ntile = (ncell_max + tile - 1) / tile; r.a.win = 0;
  r.mark(PH);
  LIN_LAUNCH(r, (relem_lin_phase_kernel<PH, NCH>), r.a.count * r.a.ntile, LIN_THREADS, smem);
  r.mark(PH);
}

// inside phase of the scanner's second pass: at most d + 2 cells per sequence contain Ys (see the kernel)
template <int PH> static void launch_phase_win(Runner& r, int d, int tile, int smem) {
  int ncell_max = r.a.lay.Lmax + 1 - d;
  if (ncell_max <= 0) return;
  if (ncell_max > d + 2) ncell_max = d + 2;
  if (tile > ncell_max) tile = ncell_max;
  tile = fit_tile(r, tile, ncell_max);
  r.a.d = d; r.a.tile = tile; r.a.ntile = (ncell_max + tile - 1) / tile; r.a.win = 1;
  r.mark(PH);
  LIN_LAUNCH(r, (relem_lin_phase_kernel<PH, 1>), r.a.count * r.a.ntile, LIN_THREADS, smem);
  r.mark(PH);
}

template <int PH, int MODE> static void launch_phase3(Runner& r, int d, int tile, int smem) {
  const int ncell_max = r.a.lay.Lmax + 1 - d;
  if (ncell_max <= 0) return;
  if (tile > ncell_max) tile = ncell_max;
  tile = fit_tile(r, tile, ncell_max);
  r.a.d = d; r.a.tile = tile; r.a.ntile = (ncell_max + tile - 1) / tile; r.a.win = 0;
  r.mark(PH);
  LIN_LAUNCH(r, (relem_lin_phase_kernel<PH, 1, MODE>), r.a.count * r.a.ntile, LIN_THREADS, smem);
  r.mark(PH);
}

// small chunks: every phase of one diagonal and direction in one launch (relem_lin_diag_kernel)
template <int DIR, int NCH> static void launch_diag(Runner& r, int d) {
  const int ncell_max = r.a.lay.Lmax + 1 - d;
  if (ncell_max <= 0) return;
  int tile = fit_tile(r, r.tile_d < ncell_max ? r.tile_d : ncell_max, ncell_max);
  r.a.d = d; r.a.tile = tile; r.a.ntile = (ncell_max + tile - 1) / tile; r.a.win = 0;
  r.mark(DIR ? 15 : 14);
  LIN_LAUNCH(r, (relem_lin_diag_kernel<DIR, NCH>), r.a.count * r.a.ntile, LIN_THREADS,
             DIR ? r.smem_out : (r.smem_in > r.smem_inb ? r.smem_in : r.smem_inb));
  r.mark(DIR ? 15 : 14);
}

// scatter mode: zero b^P and the flank part of b^L of every slot of the chunk before its outside pass
template <int NCH> static void launch_zero(Runner& r) {
#if LIN_SCATTER_ILOOP
  const unsigned long long per = (unsigned long long)NCH * r.a.lay.bch / 2;   // double2 stores per slot (the larger part)
  long long nt = (long long)((per + 8ull * LIN_THREADS - 1) / (8ull * LIN_THREADS));   // ~8 stores per thread ...
  const long long want = (long long)r.fill * r.resident_ctas;
  if ((long long)r.a.count * nt > 4 * want) nt = std::max<long long>(1, 4 * want / r.a.count);   // ... but not more CTAs than the GPU chews
  r.a.ntile = (int)nt; r.a.tile = 0; r.a.d = 0; r.a.win = 0;
  LIN_LAUNCH(r, (relem_lin_zero_kernel<NCH>), r.a.count * r.a.ntile, LIN_THREADS, 0);
#else
  (void)r;
#endif
}

template <int NCH> static void run_chunk(Runner& r, bool filter, int NT) {
  const int W = r.a.lay.Wmax, cnt = r.a.count;
  LIN_LAUNCH(r, relem_lin_prep_kernel, cnt, LIN_THREADS, r.smem_small);
  if (filter) {
    for (int d = 3; d <= W; ++d) launch_phase<PH_K0_IN, 1>(r, d, r.tile_k0, r.smem_k0);
    LIN_LAUNCH(r, (relem_lin_ext_kernel<0, 1>), cnt, 32, r.smem_small);
    LIN_LAUNCH(r, (relem_lin_ext_kernel<1, 1>), cnt, 32, r.smem_small);
    for (int d = W; d >= 3; --d) launch_phase<PH_K0_OUT, 1>(r, d, r.tile_k0, r.smem_k0);
    LIN_LAUNCH(r, relem_lin_filter_kernel, cnt, LIN_THREADS, r.smem_small);
  }
  const bool fuse = LIN_SCATTER_ILOOP && r.fuse;
  for (int d = 0; d <= W; ++d) {
    if (fuse) { launch_diag<0, NCH>(r, d); continue; }
    launch_phase<PH_IN_L, 1>(r, d, r.tile_d, r.smem_in);
    if (d >= 5) {
      launch_phase<PH_IN_P, 1>(r, d, r.tile_p, r.smem_in);
      launch_phase<PH_IN_B, 1>(r, d, r.tile_d, r.smem_inb);
    }
    if (d >= 3) launch_phase<PH_IN_E, 1>(r, d, r.tile_e, r.smem_in);
  }
  LIN_LAUNCH(r, (relem_lin_ext_kernel<2, NCH>), cnt, 32, r.smem_ext_in);
  launch_zero<NCH>(r);
  LIN_LAUNCH(r, (relem_lin_ext_kernel<3, NCH>), cnt, 32, r.smem_ext_out);
  for (int d = W; d >= 0; --d) {
    if (fuse) { launch_diag<1, NCH>(r, d); continue; }
    if (d >= 3) launch_phase<PH_OUT_EM, NCH>(r, d, r.tile_p, r.smem_out);
#if LIN_SCATTER_ILOOP
    if (d >= 3) launch_phase<PH_OUT_ES, NCH>(r, d, r.tile_q, r.smem_out);
#endif
    if (d >= 5) {
      launch_phase<PH_OUT_B, NCH>(r, d, r.tile_d, r.smem_out);
      launch_phase<PH_OUT_P, NCH>(r, d, r.tile_e, r.smem_out);
#if !LIN_SCATTER_ILOOP
      launch_phase<PH_OUT_PQ, NCH>(r, d, r.tile_q, r.smem_out);
#endif
    }
    launch_phase<PH_OUT_L, NCH>(r, d, r.tile_d, r.smem_out);
#if !LIN_SCATTER_ILOOP
    if (d >= 1 && d <= r.cmax) {
      launch_phase<PH_OUT_LF, NCH>(r, d, r.tile_f, r.smem_out);
      launch_phase<PH_OUT_LR, NCH>(r, d, r.tile_f, r.smem_out);
    }
#endif
  }
  LIN_LAUNCH(r, (relem_lin_fold_kernel<NCH>), cnt, LIN_THREADS, NCH * NT * 8 + 16);
}

// scanner: unconstrained inside/outside (start / inner posteriors, E[N]) -> Ys -> start-constrained inside/outside
// (end posteriors) -> Ye.  The Viterbi pass is launched by the caller afterwards (it must stay bit-exact).
static void run_chunk_scan(Runner& r, bool filter, int NT) {
  const int W = r.a.lay.Wmax, cnt = r.a.count;
  LIN_LAUNCH(r, relem_lin_prep_kernel, cnt, LIN_THREADS, r.smem_small);
  if (filter) {
    for (int d = 3; d <= W; ++d) launch_phase<PH_K0_IN, 1>(r, d, r.tile_k0, r.smem_k0);
    LIN_LAUNCH(r, (relem_lin_ext_kernel<0, 1>), cnt, 32, r.smem_small);
    LIN_LAUNCH(r, (relem_lin_ext_kernel<1, 1>), cnt, 32, r.smem_small);
    for (int d = W; d >= 3; --d) launch_phase<PH_K0_OUT, 1>(r, d, r.tile_k0, r.smem_k0);
    LIN_LAUNCH(r, relem_lin_filter_kernel, cnt, LIN_THREADS, r.smem_small);
  }
  for (int pass = 1; pass <= 2; ++pass) {
    for (int d = 0; d <= W; ++d) {
      if (pass == 1) {
        launch_phase<PH_IN_L, 1>(r, d, r.tile_d, r.smem_in);
        if (d >= 5) {
          launch_phase<PH_IN_P, 1>(r, d, r.tile_p, r.smem_in);
          launch_phase<PH_IN_B, 1>(r, d, r.tile_d, r.smem_inb);
        }
        if (d >= 3) launch_phase<PH_IN_E, 1>(r, d, r.tile_e, r.smem_in);
      } else {
        launch_phase_win<PH_IN_L>(r, d, r.tile_d, r.smem_in);
        if (d >= 5) {
          launch_phase_win<PH_IN_P>(r, d, r.tile_p, r.smem_in);
          launch_phase_win<PH_IN_B>(r, d, r.tile_d, r.smem_inb);
        }
        if (d >= 3) launch_phase_win<PH_IN_E>(r, d, r.tile_e, r.smem_in);
      }
    }
    launch_zero<1>(r);
    if (pass == 1) {
      LIN_LAUNCH(r, (relem_lin_ext_kernel<2, 1, 1>), cnt, 32, r.smem_ext_in);
      LIN_LAUNCH(r, (relem_lin_ext_kernel<3, 1, 1>), cnt, 32, r.smem_ext_out);
    } else {
      LIN_LAUNCH(r, (relem_lin_ext_kernel<2, 1, 2>), cnt, 32, r.smem_ext_in);
      LIN_LAUNCH(r, (relem_lin_ext_kernel<3, 1, 2>), cnt, 32, r.smem_ext_out);
    }
    for (int d = W; d >= 0; --d) {
      if (pass == 1) {
        if (d >= 3) launch_phase3<PH_OUT_EM, 1>(r, d, r.tile_p, r.smem_out);
#if LIN_SCATTER_ILOOP
        if (d >= 3) launch_phase3<PH_OUT_ES, 1>(r, d, r.tile_q, r.smem_out);
#endif
        if (d >= 5) {
          launch_phase3<PH_OUT_B, 1>(r, d, r.tile_d, r.smem_out);
          launch_phase3<PH_OUT_P, 1>(r, d, r.tile_e, r.smem_out);
#if !LIN_SCATTER_ILOOP
          launch_phase3<PH_OUT_PQ, 1>(r, d, r.tile_q, r.smem_out);
#endif
        }
        launch_phase3<PH_OUT_L, 1>(r, d, r.tile_d, r.smem_out);
#if !LIN_SCATTER_ILOOP
        if (d >= 1 && d <= r.cmax) {
          launch_phase3<PH_OUT_LF, 1>(r, d, r.tile_f, r.smem_out);
          launch_phase3<PH_OUT_LR, 1>(r, d, r.tile_f, r.smem_out);
        }
#endif
      } else {
        if (d >= 3) launch_phase3<PH_OUT_EM, 2>(r, d, r.tile_p, r.smem_out);
#if LIN_SCATTER_ILOOP
        if (d >= 3) launch_phase3<PH_OUT_ES, 2>(r, d, r.tile_q, r.smem_out);
#endif
        if (d >= 5) {
          launch_phase3<PH_OUT_B, 2>(r, d, r.tile_d, r.smem_out);
          launch_phase3<PH_OUT_P, 2>(r, d, r.tile_e, r.smem_out);
#if !LIN_SCATTER_ILOOP
          launch_phase3<PH_OUT_PQ, 2>(r, d, r.tile_q, r.smem_out);
#endif
        }
        launch_phase3<PH_OUT_L, 2>(r, d, r.tile_d, r.smem_out);
#if !LIN_SCATTER_ILOOP
        if (d >= 1 && d <= r.cmax) {
          launch_phase3<PH_OUT_LF, 2>(r, d, r.tile_f, r.smem_out);
          launch_phase3<PH_OUT_LR, 2>(r, d, r.tile_f, r.smem_out);
        }
#endif
      }
    }
    if (pass == 1) LIN_LAUNCH(r, (relem_lin_scanfold_kernel<1>), cnt, LIN_THREADS, NT * 8 + 16);
    else LIN_LAUNCH(r, (relem_lin_scanfold_kernel<2>), cnt, LIN_THREADS, 16);
  }
}

int lin_estep_launch(LinState* st, const LinLaunch& in, float* kernel_ms, int* launches, std::string& err) {
  if (kernel_ms) *kernel_ms = 0.f;
  if (launches) *launches = 0;
  LinConst hc;
  hc.h = in.h; hc.p = in.p; hc.en = in.en; hc.el = in.el; hc.k0 = in.kappa0; hc.k0sq = in.kappa0 * in.kappa0;
  Runner r;
  LinKArgs& a = r.a;
  a.b = in.b; a.out = in.out; a.flag = in.flag;
  const int nseq = in.b.nseq;
  a.lay = make_lin_layout(std::max(1, in.Lmax), in.max_span, in.h, in.nch);
  const LinLayout& lay = a.lay;
  r.smem_small = lay.sm_warp;
  r.smem_k0 = lay.sm_warp + LIN_WARPS * 128 * 4;
  r.smem_in = lay.sm_warp + LIN_WARPS * lay.warp_bytes_in;
  r.smem_inb = lay.sm_warp + LIN_WARPS * lay.warp_bytes_inb;
  r.smem_out = lay.sm_warp + LIN_WARPS * lay.warp_bytes_out;
  r.smem_ext_in = lay.sm_warp + lay.warp_bytes_in;
  r.smem_ext_out = lay.sm_warp + lay.warp_bytes_out;
  r.cmax = std::min(30, std::min((int)in.en.max_iloop, lay.Wmax - 7));
  if (in.sm_count > 0) r.resident_ctas = in.sm_count * 7;
  const int NT = in.p.n_theta;
  if (const char* e = std::getenv("RELEM_FILL")) r.fill = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("RELEM_TILE_K0")) r.tile_k0 = std::max(4, std::atoi(e));
  if (const char* e = std::getenv("RELEM_TILE_D")) r.tile_d = std::max(4, std::atoi(e));
  if (const char* e = std::getenv("RELEM_TILE_P")) r.tile_p = std::max(4, std::atoi(e));
  if (const char* e = std::getenv("RELEM_TILE_E")) r.tile_e = std::max(4, std::atoi(e));
  if (const char* e = std::getenv("RELEM_TILE_F")) r.tile_f = std::max(4, std::atoi(e));
  if (const char* e = std::getenv("RELEM_TILE_Q")) r.tile_q = std::max(4, std::atoi(e));
  // kappa0 powers [0, KP) followed by the separable interior-loop table G[32][32]; the device fills G (gtab kernel)
  const int KP = std::max(256, (lay.Wmax + 3 + 31) & ~31);
  std::vector<double> kp(KP + 1024, 0.);
  for (int t = 0; t < KP; ++t) kp[t] = std::pow(in.kappa0, (double)t);
  size_t per = (size_t)lay.stride * sizeof(double);
#ifdef RELEM_HOST_EMU
  LC = hc;
  // chunks of up to 4 sequences share one launch sequence, like a (tiny) device chunk: sequences of different length
  // in one chunk exercise everything that depends on the chunk-wide layout versus the per-sequence one
  const int emu_chunk = 4;
  if (per * emu_chunk > st->scratch_bytes) {
    std::free(st->scratch);
    st->scratch = std::malloc(per * emu_chunk);
    st->scratch_bytes = st->scratch ? per * emu_chunk : 0;
  }
  if (!st->scratch) { err = "scratch allocation failed"; return 3; }
  std::free(st->k0pow);
  st->k0pow = std::malloc(kp.size() * 8);
  std::memcpy(st->k0pow, kp.data(), kp.size() * 8);
  a.scratch = (double*)st->scratch; a.k0pow = (const double*)st->k0pow; a.kp_n = KP;
  r.smem.assign(std::max(std::max(r.smem_out, std::max(r.smem_in, r.smem_inb)), NT * in.nch * 8 + 16) + 64, 0);
  LIN_LAUNCH(r, relem_lin_gtab_kernel, 1, LIN_THREADS, 0);
  for (int k = 0; k < nseq; k += emu_chunk) {
    // poison: the gather passes must never read an entry they did not write
    { double* p = (double*)st->scratch; for (size_t z = 0; z < per * emu_chunk / 8; ++z) p[z] = std::nan(""); }
    a.base = k; a.count = std::min(emu_chunk, nseq - k);
    if (const char* e = std::getenv("RELEM_FUSE_CELLS")) r.fuse = (long long)a.count * a.lay.Lmax <= std::atoll(e);
    if (in.nch == 2) run_chunk<2>(r, in.en.filter != 0, NT);
    else run_chunk<1>(r, in.en.filter != 0, NT);
  }
  if (launches) *launches = r.launches;
  return 0;
#else
  if (r.smem_out > 227 * 1024) { err = "pattern too large for the linear-space kernel's shared memory"; return 1; }
  // a scratch buffer that already holds the whole batch needs no new sizing (cudaMemGetInfo costs milliseconds,
  // which matters for a training minibatch of ~128 sequences)
  long long by_mem = (long long)(st->scratch_bytes / per);
  if (by_mem < nseq) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    by_mem = (long long)(((double)(free_b + st->scratch_bytes) * 0.70) / (double)per);
  }
  if (in.max_slots > 0) by_mem = std::min<long long>(by_mem, in.max_slots);
  if (by_mem < 1) { err = "not enough device memory for one sequence slot"; return 3; }
  // two lanes (streams) with half of the slots each: the dependent launches of one lane fill the drain / ramp-up gaps
  // of the other.  Measured (gpurun_out/bench62_*): +11 % at 128 sequences, +6 % at 512 and 2 048, +2.5 % at 20 000;
  // four lanes are no better.  Not when memory forces chunks so small that halving them would starve a launch.
  int nlanes = ((nseq >= 64 && by_mem >= nseq) || (nseq >= 4096 && by_mem >= 2048)) ? 2 : 1;
  if (const char* e = std::getenv("RELEM_LANES")) nlanes = std::max(1, std::min(4, std::atoi(e)));
  r.st = st;
  if (const char* e = std::getenv("RELEM_PHASE_TIMING")) r.timing = std::atoi(e) != 0;
  if (r.timing) nlanes = 1;   // per-launch events only mean something when nothing else shares the GPU
  if (nlanes > by_mem) nlanes = (int)by_mem;   // every lane needs at least one slot
  long long nslots = std::max<long long>(1, std::min<long long>((nseq + nlanes - 1) / nlanes, by_mem / nlanes));
  // keep an existing scratch buffer when it is close to what we would ask for (free memory fluctuates a little
  // from call to call; re-allocating ~100 GB costs more than a slightly smaller chunk)
  const long long have = (long long)(st->scratch_bytes / ((size_t)nlanes * per));
  if (have >= 1 && have < nslots && have * 10 >= nslots * 8) nslots = have;
  size_t need = (size_t)nslots * nlanes * per;
  if (need > st->scratch_bytes) {
    if (st->scratch) cudaFree(st->scratch);
    st->scratch = nullptr; st->scratch_bytes = 0;
    if (cudaMalloc(&st->scratch, need) != cudaSuccess) { err = "scratch allocation failed"; return 3; }
    st->scratch_bytes = need;
  }
  if (st->k0pow_n < kp.size()) {
    if (st->k0pow) cudaFree(st->k0pow);
    st->k0pow = nullptr; st->k0pow_n = 0;
    if (cudaMalloc(&st->k0pow, kp.size() * 8) != cudaSuccess) { err = "allocation failed"; return 3; }
    st->k0pow_n = kp.size();
  }
  r.stream = (cudaStream_t)in.stream;
  cudaError_t e = cudaSuccess;
  const bool new_kp = st->k0pow_kappa != in.kappa0 || st->k0pow_kp != KP;
  if (new_kp) e = cudaMemcpyAsync(st->k0pow, kp.data(), kp.size() * 8, cudaMemcpyHostToDevice, r.stream);
  if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(LC, &hc, sizeof(LinConst), 0, cudaMemcpyHostToDevice, r.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(r.stream);  // kp / hc are stack objects
  if (e != cudaSuccess) { err = std::string("constant upload: ") + cudaGetErrorString(e); return 2; }
  a.scratch = (double*)st->scratch; a.k0pow = (const double*)st->k0pow; a.kp_n = KP;
  cudaStream_t main_stream = r.stream;
  if (!st->ev0) { cudaEventCreate(&st->ev0); cudaEventCreate(&st->ev1); }
  cudaEvent_t e0 = st->ev0, e1 = st->ev1;
  // the G table depends on the energy tables and kappa0 only; energy parameters change rarely but cheaply: refill
  LIN_LAUNCH(r, relem_lin_gtab_kernel, 1, LIN_THREADS, 0);
  st->k0pow_kappa = in.kappa0; st->k0pow_kp = KP;
  cudaEventRecord(e0, main_stream);
  if (nlanes >= 2) {
    for (int k = 0; k < nlanes; ++k) {
      if (!st->lane[k]) cudaStreamCreateWithFlags(&st->lane[k], cudaStreamNonBlocking);
      if (!st->lane_done[k]) cudaEventCreateWithFlags(&st->lane_done[k], cudaEventDisableTiming);
      cudaStreamWaitEvent(st->lane[k], e0, 0);
    }
  }
  // chunks of at most this many sequence positions run one fused launch per diagonal (see relem_lin_diag_kernel)
  long long fuse_cells = LIN_FUSE_CELLS;
  if (const char* e = std::getenv("RELEM_FUSE_CELLS")) fuse_cells = std::atoll(e);
  int chunk = 0;
  for (int base = 0; base < nseq; base += (int)nslots, ++chunk) {
    const int ln = chunk % nlanes;
    r.stream = nlanes >= 2 ? st->lane[ln] : main_stream;
    a.scratch = (double*)st->scratch + (size_t)ln * (size_t)nslots * lay.stride;
    a.base = base; a.count = std::min<int>((int)nslots, nseq - base);
    r.fuse = (long long)a.count * lay.Lmax <= fuse_cells;
    if (in.nch == 2) run_chunk<2>(r, in.en.filter != 0, NT);
    else run_chunk<1>(r, in.en.filter != 0, NT);
  }
  if (nlanes >= 2) {
    for (int k = 0; k < nlanes; ++k) {
      cudaEventRecord(st->lane_done[k], st->lane[k]);
      cudaStreamWaitEvent(main_stream, st->lane_done[k], 0);
    }
  }
  r.stream = main_stream;
  e = cudaGetLastError();
  cudaEventRecord(e1, r.stream);
  if (e != cudaSuccess) { err = std::string("linear-space kernel launch: ") + cudaGetErrorString(e); return 2; }
  e = cudaEventSynchronize(e1);
  if (e != cudaSuccess) { err = std::string("linear-space kernels: ") + cudaGetErrorString(e); return 2; }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (kernel_ms) *kernel_ms = ms;
  if (launches) *launches = r.launches;
  for (int k = 0; k < LIN_NPHASE_SLOTS; ++k) { st->phase_ms[k] = 0.f; st->phase_launches[k] = 0; }
  for (size_t k = 0; k < r.marks.size(); ++k) {
    float pm = 0.f;
    cudaEventElapsedTime(&pm, st->ev_pool[2 * k], st->ev_pool[2 * k + 1]);
    st->phase_ms[r.marks[k]] += pm;
    st->phase_launches[r.marks[k]] += 1;
  }
  return 0;
#endif
}

int lin_phase_timing(const LinState* st, const char** names, float* ms, int* launches, int cap) {
  static const char* kName[LIN_NPHASE_SLOTS] = {
      "relem_lin_phase_kernel<0> K0 inside", "relem_lin_phase_kernel<1> K0 outside", "relem_lin_phase_kernel<2> inside L",
      "relem_lin_phase_kernel<3> inside P", "relem_lin_phase_kernel<4> inside B,2,1,M", "relem_lin_phase_kernel<5> inside E",
      "relem_lin_phase_kernel<6> outside E,M", "relem_lin_phase_kernel<7> outside 1,B,2",
      "relem_lin_phase_kernel<8> outside P (2<-P, stack, exterior)", "relem_lin_phase_kernel<9> outside L (hairpin, parent)",
      "relem_lin_phase_kernel<10> outside L left flanks", "relem_lin_phase_kernel<11> outside L right flanks",
      "relem_lin_phase_kernel<12> outside P enclosing interior loops",
      "relem_lin_phase_kernel<13> outside E interior loops (scatter to P and flanks)",
      "relem_lin_diag_kernel<0> inside, all phases of a diagonal (small chunks)",
      "relem_lin_diag_kernel<1> outside, all phases of a diagonal (small chunks)"};
  int n = 0;
  if (!st) return 0;
  for (int k = 0; k < LIN_NPHASE_SLOTS && n < cap; ++k) {
    if (!st->phase_launches[k]) continue;
    names[n] = kName[k]; ms[n] = st->phase_ms[k]; launches[n] = st->phase_launches[k];
    ++n;
  }
  return n;
}

int lin_scan_launch(LinState* st, const LinScanLaunch& in, lin_chunk_fn after_chunk, void* user, float* kernel_ms,
                    int* launches, std::string& err) {
  if (kernel_ms) *kernel_ms = 0.f;
  if (launches) *launches = 0;
  LinConst hc;
  hc.h = in.h; hc.p = in.p; hc.en = in.en; hc.el = in.el; hc.k0 = in.kappa0; hc.k0sq = in.kappa0 * in.kappa0;
  Runner r;
  LinKArgs& a = r.a;
  std::memset(&a.out, 0, sizeof(a.out));
  a.b = in.b; a.so = in.so; a.flag = in.flag;
  const int nseq = in.b.nseq;
  a.lay = make_lin_layout(std::max(1, in.Lmax), in.max_span, in.h, 1);
  const LinLayout& lay = a.lay;
  r.smem_small = lay.sm_warp;
  r.smem_k0 = lay.sm_warp + LIN_WARPS * 128 * 4;
  r.smem_in = lay.sm_warp + LIN_WARPS * lay.warp_bytes_in;
  r.smem_inb = lay.sm_warp + LIN_WARPS * lay.warp_bytes_inb;
  r.smem_out = lay.sm_warp + LIN_WARPS * lay.warp_bytes_out;
  r.smem_ext_in = lay.sm_warp + lay.warp_bytes_in;
  r.smem_ext_out = lay.sm_warp + lay.warp_bytes_out;
  r.cmax = std::min(30, std::min((int)in.en.max_iloop, lay.Wmax - 7));
  const int NT = in.p.n_theta;
  const int KP = std::max(256, (lay.Wmax + 3 + 31) & ~31);
  std::vector<double> kp(KP + 1024, 0.);
  for (int t = 0; t < KP; ++t) kp[t] = std::pow(in.kappa0, (double)t);
  size_t per = (size_t)lay.stride * sizeof(double);
  LinChunkView cv;
  cv.stride = lay.stride; cv.masks_off = lay.masks; cv.mask_words = lay.mask_words;
#ifdef RELEM_HOST_EMU
  LC = hc;
  const int emu_chunk = 4;   // see lin_estep_launch
  if (per * emu_chunk > st->scratch_bytes) {
    std::free(st->scratch);
    st->scratch = std::malloc(per * emu_chunk);
    st->scratch_bytes = st->scratch ? per * emu_chunk : 0;
  }
  if (!st->scratch) { err = "scratch allocation failed"; return 3; }
  std::free(st->k0pow);
  st->k0pow = std::malloc(kp.size() * 8);
  std::memcpy(st->k0pow, kp.data(), kp.size() * 8);
  a.scratch = (double*)st->scratch; a.k0pow = (const double*)st->k0pow; a.kp_n = KP;
  r.smem.assign(std::max(std::max(r.smem_out, r.smem_inb), NT * 8 + 16) + 64, 0);
  LIN_LAUNCH(r, relem_lin_gtab_kernel, 1, LIN_THREADS, 0);
  for (int k = 0; k < nseq; k += emu_chunk) {
    { double* p = (double*)st->scratch; for (size_t z = 0; z < per * emu_chunk / 8; ++z) p[z] = std::nan(""); }
    a.base = k; a.count = std::min(emu_chunk, nseq - k);
    run_chunk_scan(r, in.en.filter != 0, NT);
    cv.base = k; cv.count = a.count; cv.scratch = a.scratch; cv.stream = nullptr;
    if (after_chunk && after_chunk(user, cv)) { err = "Viterbi launch failed"; return 2; }
  }
  if (launches) *launches = r.launches;
  return 0;
#else
  if (r.smem_out > 227 * 1024) { err = "pattern too large for the linear-space kernel's shared memory"; return 1; }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  long long by_mem = (long long)(((double)(free_b + st->scratch_bytes) * 0.70) / (double)per);
  if (in.max_slots > 0) by_mem = std::min<long long>(by_mem, in.max_slots);
  if (by_mem < 1) { err = "not enough device memory for one sequence slot"; return 3; }
  // two lanes as in lin_estep_launch: while one lane's Viterbi kernel (few, fat CTAs) runs, the other lane's phase
  // kernels fill the SMs.  The Viterbi launches themselves are serialised by the caller (they share one value table).
  int nlanes = (nseq >= 64 && by_mem >= 2) ? 2 : 1;
  if (const char* e = std::getenv("RELEM_SCAN_LANES")) nlanes = std::max(1, std::min(4, std::atoi(e)));
  if (nlanes > by_mem) nlanes = (int)by_mem;
  long long nslots = std::max<long long>(1, std::min<long long>((nseq + nlanes - 1) / nlanes, by_mem / nlanes));
  const long long have = (long long)(st->scratch_bytes / ((size_t)nlanes * per));
  if (have >= 1 && have < nslots && have * 10 >= nslots * 8) nslots = have;
  size_t need = (size_t)nslots * nlanes * per;
  if (need > st->scratch_bytes) {
    if (st->scratch) cudaFree(st->scratch);
    st->scratch = nullptr; st->scratch_bytes = 0;
    if (cudaMalloc(&st->scratch, need) != cudaSuccess) { err = "scratch allocation failed"; return 3; }
    st->scratch_bytes = need;
  }
  if (st->k0pow_n < kp.size()) {
    if (st->k0pow) cudaFree(st->k0pow);
    st->k0pow = nullptr; st->k0pow_n = 0;
    if (cudaMalloc(&st->k0pow, kp.size() * 8) != cudaSuccess) { err = "allocation failed"; return 3; }
    st->k0pow_n = kp.size();
  }
  r.stream = (cudaStream_t)in.stream;
  cudaError_t e = cudaMemcpyAsync(st->k0pow, kp.data(), kp.size() * 8, cudaMemcpyHostToDevice, r.stream);
  if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(LC, &hc, sizeof(LinConst), 0, cudaMemcpyHostToDevice, r.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(r.stream);
  if (e != cudaSuccess) { err = std::string("constant upload: ") + cudaGetErrorString(e); return 2; }
  a.scratch = (double*)st->scratch; a.k0pow = (const double*)st->k0pow; a.kp_n = KP;
  if (!st->ev0) { cudaEventCreate(&st->ev0); cudaEventCreate(&st->ev1); }
  cudaEvent_t e0 = st->ev0, e1 = st->ev1;   // owned by the state: nothing to release on the error returns below
  cudaStream_t main_stream = r.stream;
  LIN_LAUNCH(r, relem_lin_gtab_kernel, 1, LIN_THREADS, 0);
  cudaEventRecord(e0, main_stream);
  if (nlanes >= 2) {
    for (int k = 0; k < nlanes; ++k) {
      if (!st->lane[k]) cudaStreamCreateWithFlags(&st->lane[k], cudaStreamNonBlocking);
      if (!st->lane_done[k]) cudaEventCreateWithFlags(&st->lane_done[k], cudaEventDisableTiming);
      cudaStreamWaitEvent(st->lane[k], e0, 0);
    }
  }
  int chunk = 0;
  for (int base = 0; base < nseq; base += (int)nslots, ++chunk) {
    const int ln = chunk % nlanes;
    r.stream = nlanes >= 2 ? st->lane[ln] : main_stream;
    a.scratch = (double*)st->scratch + (size_t)ln * (size_t)nslots * lay.stride;
    a.base = base; a.count = std::min<int>((int)nslots, nseq - base);
    run_chunk_scan(r, in.en.filter != 0, NT);
    cv.base = a.base; cv.count = a.count; cv.scratch = a.scratch; cv.stream = (void*)r.stream;
    if (after_chunk && after_chunk(user, cv)) { err = "Viterbi launch failed"; return 2; }
  }
  if (nlanes >= 2) {
    for (int k = 0; k < nlanes; ++k) {
      cudaEventRecord(st->lane_done[k], st->lane[k]);
      cudaStreamWaitEvent(main_stream, st->lane_done[k], 0);
    }
  }
  r.stream = main_stream;
  e = cudaGetLastError();
  cudaEventRecord(e1, r.stream);
  if (e != cudaSuccess) { err = std::string("linear-space kernel launch: ") + cudaGetErrorString(e); return 2; }
  e = cudaEventSynchronize(e1);
  if (e != cudaSuccess) { err = std::string("linear-space kernels: ") + cudaGetErrorString(e); return 2; }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (kernel_ms) *kernel_ms = ms;
  if (launches) *launches = r.launches;
  return 0;
#endif
}

// ------------------------------------------------------------------------------------------------ fp64 peaks
#ifndef RELEM_HOST_EMU
// 8 independent dependent-FMA chains per thread: enough instruction-level parallelism to keep the fp64 pipe full
__global__ void __launch_bounds__(256) relem_fp64_dfma_kernel(double* sink, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
  for (int k = 0; k < iters; ++k) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double v = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (v == 12345.678) sink[0] = v;   // never true: keeps the loop alive
}
__global__ void __launch_bounds__(256) relem_fp64_exp_kernel(double* sink, int iters, double a) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + .25, x2 = x0 + .5, x3 = x0 + .75;
  for (int k = 0; k < iters; ++k) {
    x0 = exp(x0 * a); x1 = exp(x1 * a); x2 = exp(x2 * a); x3 = exp(x3 * a);
  }
  double v = (x0 + x1) + (x2 + x3);
  if (v == 12345.678) sink[0] = v;
}
#endif

int lin_fp64_peak(void* stream, int sm_count, double* dfma_per_s, double* exp_per_s, std::string& err) {
#ifdef RELEM_HOST_EMU
  (void)stream; (void)sm_count; (void)dfma_per_s; (void)exp_per_s;
  err = "no micro-benchmark in the emulation";
  return 1;
#else
  cudaStream_t st = (cudaStream_t)stream;
  double* sink = nullptr;
  if (cudaMalloc(&sink, 8) != cudaSuccess) { err = "allocation failed"; return 3; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = std::max(1, sm_count) * 8, threads = 256;
  float ms = 0.f;
  double best_f = 0., best_e = 0.;
  for (int rep = 0; rep < 4; ++rep) {   // first repetition warms up
    const int it_f = 20000, it_e = 1000;
    cudaEventRecord(e0, st);
    relem_fp64_dfma_kernel<<<grid, threads, 0, st>>>(sink, it_f, 0.999999, 1e-6);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) best_f = std::max(best_f, (double)grid * threads * 8. * it_f / (ms * 1e-3));
    cudaEventRecord(e0, st);
    relem_fp64_exp_kernel<<<grid, threads, 0, st>>>(sink, it_e, 0.5);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) best_e = std::max(best_e, (double)grid * threads * 4. * it_e / (ms * 1e-3));
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  if (e != cudaSuccess) { err = cudaGetErrorString(e); return 2; }
  if (dfma_per_s) *dfma_per_s = best_f;
  if (exp_per_s) *exp_per_s = best_e;
  return 0;
#endif
}

}  // namespace lin
}  // namespace relem
