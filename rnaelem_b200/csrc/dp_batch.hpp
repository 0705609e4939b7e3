// dp_batch.hpp -- batch / result views shared by the log-space kernels (dp_kernels.cuh) and the linear-space
// kernels (relem_lin.cu).
#ifndef RELEM_DP_BATCH_HPP
#define RELEM_DP_BATCH_HPP

namespace relem {
namespace dp {

struct BatchView {
  int nseq;
  const unsigned char* seq;   // concatenated base codes
  const long long* off;       // [nseq+1]
  const double* ws;           // concatenated position weights
  const unsigned char* kind;  // [nseq]
  const int* order;           // processing order (longest first)
};

struct EstepOut {   // device arrays, per sequence
  double* Z;        // [nseq][3]
  double* ENo;      // [nseq][n_theta]
  double* ENx;
  double* EH;       // [nseq][4]
  double* bpp_eff;  // [nseq]
  unsigned char* skipped;
  unsigned long long* prof;  // [16] cycles per phase summed over CTAs (thread 0 clocks), may be null
};

// device arrays of the scanner's results (sequence n at off[n]; PyeL at off[n]+n)
struct ScanOut {
  double* PysL; double* PyeL; double* PyiL;
  int* psihat; char* rss; int* Ys; int* Ye; double* exist; double* EN /*[nseq][n_theta]*/; double* ZL;
};

}  // namespace dp
}  // namespace relem
#endif
