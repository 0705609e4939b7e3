// lin_api.hpp -- host-side interface between relem_api.cu (the C ABI) and relem_lin.cu (the scaled linear-space
// E-step kernels, dp_lin.cuh).  relem_lin.cu is its own translation unit so that it can be compiled with FMA
// contraction on, while the Viterbi code of relem_api.cu keeps -fmad=false.
#ifndef RELEM_LIN_API_HPP
#define RELEM_LIN_API_HPP
#include <string>

#include "dp_batch.hpp"
#include "dp_common.cuh"
#include "lin_model.hpp"

namespace relem {
namespace lin {

struct LinState;  // owns the per-slot scratch of the linear-space kernels
LinState* lin_state_create();
void lin_state_destroy(LinState*);

struct LinLaunch {
  LinHMM h;
  LinParams p;
  dp::DevEnergy en, el;
  double kappa0;          // per-base scale of the energy-only filter pass
  dp::BatchView b;
  int Lmax, max_span;
  dp::EstepOut out;       // NCH = 1: ENo receives ENo-ENx and ENx zeros (likewise EH)
  unsigned char* flag;    // [nseq] device: 1 = sequence left the fp64 range, re-run it on the log-space path
  int nch;                // 1 = difference only, 2 = both boundary conditions
  int sm_count;
  void* stream;           // cudaStream_t
  int max_slots;          // 0 = no limit
};
// returns 0 on success; on failure err holds the message.  launches = kernels launched.
int lin_estep_launch(LinState*, const LinLaunch&, float* kernel_ms, int* launches, std::string& err);


// ---- scanner (RNAelemScanDP::calc_motif_positions, motif_scanner.hpp:186-214) on the linear-space kernels:
// unconstrained inside/outside (start / inner posteriors, E[N], Ys) and start-constrained inside/outside (end
// posteriors, Ye).  After the kernels of a chunk are enqueued `after_chunk` is called so that the caller can enqueue
// the (bit-exact, log-space) Viterbi kernel for the same sequences on the same stream while the chunk's masks are
// still in the slots.
struct LinChunkView {
  int base, count;
  const double* scratch;
  unsigned long long stride, masks_off;
  int mask_words;
  void* stream;
};
typedef int (*lin_chunk_fn)(void* user, const LinChunkView&);
struct LinScanLaunch {
  LinHMM h;
  LinParams p;
  dp::DevEnergy en, el;
  double kappa0;
  dp::BatchView b;
  int Lmax, max_span;
  dp::ScanOut so;
  unsigned char* flag;
  void* stream;
  int max_slots;
};
int lin_scan_launch(LinState*, const LinScanLaunch&, lin_chunk_fn after_chunk, void* user, float* kernel_ms, int* launches,
                    std::string& err);

// per-phase device time of the last lin_estep_launch that ran under RELEM_PHASE_TIMING=1 (static names)
int lin_phase_timing(const LinState*, const char** names, float* ms, int* launches, int cap);

// fp64 micro-benchmark (relem_fp64_peak): fused multiply-adds / exp() per second on the given stream's device
int lin_fp64_peak(void* stream, int sm_count, double* dfma_per_s, double* exp_per_s, std::string& err);

}  // namespace lin
}  // namespace relem
#endif
