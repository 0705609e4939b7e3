// dp_layout.hpp -- host-side sizing of the per-slot HBM scratch and of the CTA's shared memory.
#ifndef RELEM_DP_LAYOUT_HPP
#define RELEM_DP_LAYOUT_HPP
#include "dp_kernels.cuh"

namespace relem {
namespace dp {

// nch: posterior channels kept at once (2 for the E-step, 1 for scan / bpp); with_coupled = false sizes the
// scratch for the energy-only filter alone.
inline SlotLayout make_layout(int Lmax, int max_span, int S, int M, int n_theta, int nch, bool with_coupled,
                              int n_list_max = 0, int vit_reads_per_cta = 0) {
  SlotLayout lay;
  std::memset(&lay, 0, sizeof(lay));
  int Wmax = Lmax < max_span ? Lmax : max_span;
  lay.Lmax = Lmax; lay.Wmax = Wmax; lay.mw = (Wmax + 1 + 31) / 32;
  unsigned long long cells = (unsigned long long)(Lmax + 1) * (Wmax + 1);
  unsigned long long band = with_coupled ? NPLANE * cells * S : 0;
  unsigned long long ext = (unsigned long long)(Lmax + 1) * S;
  unsigned long long o = 0;
  auto take = [&](unsigned long long n) { unsigned long long r = o; o += (n + 1) & ~1ull; return r; };
  // vit_reads_per_cta < 0: slots of the Viterbi kernel only (value table, exterior row, emission tables, stack)
  const bool vit_only = vit_reads_per_cta < 0;
  if (vit_only) vit_reads_per_cta = -vit_reads_per_cta - 1;   // -1 -> no limit, -(k+1) -> at most k sequences per CTA
  lay.tabA = take(band);
  lay.Q0 = take(vit_only ? 0 : band);
  lay.Q1 = take(nch > 1 && !vit_only ? band : 0);
  lay.tab0 = take(vit_only ? 0 : NPLANE * cells);
  lay.q0 = take(vit_only ? 0 : NPLANE * cells);
  lay.otab = take(ext);
  lay.QO0 = take(vit_only ? 0 : ext);
  lay.QO1 = take(nch > 1 && !vit_only ? ext : 0);
  lay.otab0 = take(vit_only ? 0 : Lmax + 1);
  lay.QO00 = take(vit_only ? 0 : Lmax + 1);
  lay.emit0 = take((unsigned long long)M * Lmax);
  lay.emitT = take((unsigned long long)M * Lmax);
  lay.zeros = take(vit_only ? 0 : Lmax + 1);
  lay.G = take(vit_only ? 0 : (unsigned long long)nch * M * Lmax);
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
  // the Viterbi kernel (dp_vit.cuh) carves shared memory on its own: R sequences per CTA, as many as fit next to the
  // warp slices in ~100 KB (so that two CTAs still share an SM), at most one per warp
  {
    const int nwarp = RELEM_VIT_THREADS / 32;
    const int mask_b = (mask_bytes + 15) & ~15, row_b = (Lmax + 2 + 15) & ~15;
    lay.vit_seq_bytes = 4 * row_b + 2 * mask_b;
    lay.vit_n_max = n_list_max > S ? n_list_max : S;
    lay.vit_warp_bytes = vit_warp_bytes(S, Wmax, lay.vit_n_max);
    const int fixed = 16 + nwarp * lay.vit_warp_bytes;
    int R = (100 * 1024 - fixed - 64) / (lay.vit_seq_bytes + (int)sizeof(VitSeq) + 16);
    if (R > nwarp / 2) R = nwarp / 2;   // measured: 4 per CTA = 8 per CTA = +4 % over one (profiles/r2_viterbi.md); fewer slots
    if (R < 1) R = 1;
    if (vit_reads_per_cta > 0 && R > vit_reads_per_cta) R = vit_reads_per_cta;
    lay.vit_R = R;
    int bv = 0;
    auto smv = [&](int bytes) { int r = bv; bv += (bytes + 15) & ~15; return r; };
    lay.sm_vit_claim = smv(16);
    lay.sm_vit_ctx = smv(R * (int)sizeof(VitSeq));
    lay.sm_vit_seq = smv(R * lay.vit_seq_bytes);
    lay.sm_vit_warp = smv(lay.vit_warp_bytes * nwarp);
    lay.sm_total_vit = bv;
  }
  return lay;
}

}  // namespace dp
}  // namespace relem
#endif
