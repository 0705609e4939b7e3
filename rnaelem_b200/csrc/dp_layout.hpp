// dp_layout.hpp -- host-side sizing of the per-slot HBM scratch and of the CTA's shared memory.
#ifndef RELEM_DP_LAYOUT_HPP
#define RELEM_DP_LAYOUT_HPP
#include "dp_kernels.cuh"

namespace relem {
namespace dp {

// nch: posterior channels kept at once (2 for the E-step, 1 for scan / bpp); with_coupled = false sizes the
// scratch for the energy-only filter alone.
inline SlotLayout make_layout(int Lmax, int max_span, int S, int M, int n_theta, int nch, bool with_coupled,
                              int n_list_max = 0) {
  SlotLayout lay;
  std::memset(&lay, 0, sizeof(lay));
  int Wmax = Lmax < max_span ? Lmax : max_span;
  lay.Lmax = Lmax; lay.Wmax = Wmax; lay.mw = (Wmax + 1 + 31) / 32;
  unsigned long long cells = (unsigned long long)(Lmax + 1) * (Wmax + 1);
  unsigned long long band = with_coupled ? NPLANE * cells * S : 0;
  unsigned long long ext = (unsigned long long)(Lmax + 1) * S;
  unsigned long long o = 0;
  auto take = [&](unsigned long long n) { unsigned long long r = o; o += (n + 1) & ~1ull; return r; };
  lay.tabA = take(band);
  lay.Q0 = take(band);
  lay.Q1 = take(nch > 1 ? band : 0);
  lay.tab0 = take(NPLANE * cells);
  lay.q0 = take(NPLANE * cells);
  lay.otab = take(ext);
  lay.QO0 = take(ext);
  lay.QO1 = take(nch > 1 ? ext : 0);
  lay.otab0 = take(Lmax + 1);
  lay.QO00 = take(Lmax + 1);
  lay.emit0 = take((unsigned long long)M * Lmax);
  lay.emitT = take((unsigned long long)M * Lmax);
  lay.zeros = take(Lmax + 1);
  lay.G = take((unsigned long long)nch * M * Lmax);
  lay.stack = take(2ull * (4 * Lmax + 16));  // 4 ints per entry, depth <= 4L+16
  lay.stride = o;
  int b = 0;
  auto sm = [&](int bytes) { int r = b; b += (bytes + 15) & ~15; return r; };
  lay.sm_x = sm(Lmax + 1);
  lay.sm_sp3 = sm(Lmax + 1); lay.sm_sp4 = sm(Lmax + 1); lay.sm_sp6 = sm(Lmax + 1);
  int mask_bytes = (Lmax + 1) * lay.mw * 4;
  lay.sm_bp = sm(mask_bytes); lay.sm_lf = sm(mask_bytes); lay.sm_bp2 = sm(mask_bytes);
  lay.sm_en = sm(2 * n_theta * 8 + 8);
  lay.sm_pys = sm((Lmax + 1) * 8); lay.sm_pyi = sm((Lmax + 1) * 8); lay.sm_pye = sm((Lmax + 2) * 8);
  lay.sm_red = sm(64 * 8);
  lay.sm_ctr = sm(16);
  lay.warp_bytes = warp_sm_bytes(S, Wmax);
  lay.sm_warp = sm(lay.warp_bytes * (RELEM_CTA_THREADS / 32));
  lay.sm_total = b;
  // the Viterbi kernel (dp_vit.cuh) has its own tail: a per-diagonal work counter and one VitWarp slice per warp, placed
  // where the log-space kernels keep their warp slices
  int bv = lay.sm_warp;
  auto smv = [&](int bytes) { int r = bv; bv += (bytes + 15) & ~15; return r; };
  lay.sm_vit_ctr = smv((Wmax + 2) * 4);
  lay.vit_n_max = n_list_max > S ? n_list_max : S;
  lay.vit_warp_bytes = vit_warp_bytes(S, Wmax, lay.vit_n_max);
  lay.sm_vit_warp = smv(lay.vit_warp_bytes * (RELEM_VIT_THREADS / 32));
  lay.sm_total_vit = bv;
  return lay;
}

}  // namespace dp
}  // namespace relem
#endif
