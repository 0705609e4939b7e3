// relem_api.cu -- implementation of the C ABI declared in include/relem.h.
//
// Built by nvcc for sm_100a into rnaelem_b200/librelem.so (the product).  The same file builds, with
// -DRELEM_HOST_EMU under g++, into tests/emu/librelem_emu.so: a single-threaded host emulation of the kernel
// source used only to debug the DP logic in a container without a GPU.  The product library contains no host
// implementation of the DP and fails loudly (relem_create -> RELEM_ECUDA) when no GPU is usable.
#include "../../include/relem.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#include "dp_layout.hpp"
#include "host_model.hpp"
#include "lin_api.hpp"

#ifndef RELEM_HOST_EMU
#include <cuda_runtime.h>
#include <dlfcn.h>
#endif

using namespace relem;
using namespace relem::dp;

namespace {

std::string g_create_error;

// ---------------------------------------------------------------------------------- device abstraction
#ifdef RELEM_HOST_EMU
struct Dev {
  static bool alloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p != nullptr; }
  static void free(void* p) { std::free(p); }
  static bool h2d(void* d, const void* h, size_t n) { if (n) std::memcpy(d, h, n); return true; }
  static bool d2h(void* h, const void* d, size_t n) { if (n) std::memcpy(h, d, n); return true; }
  static bool zero(void* d, size_t n) { if (n) std::memset(d, 0, n); return true; }
};
#else
struct Dev {
  static bool alloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1) == cudaSuccess; }
  static void free(void* p) { cudaFree(p); }
  static bool h2d(void* d, const void* h, size_t n) { return !n || cudaMemcpy(d, h, n, cudaMemcpyHostToDevice) == cudaSuccess; }
  static bool d2h(void* h, const void* d, size_t n) { return !n || cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost) == cudaSuccess; }
  static bool zero(void* d, size_t n) { return !n || cudaMemset(d, 0, n) == cudaSuccess; }
};
#endif

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool reserve(size_t n) {
    if (n <= bytes && p) return true;
    if (p) Dev::free(p);
    p = nullptr; bytes = 0;
    if (!Dev::alloc(&p, n)) { p = nullptr; return false; }
    bytes = n;
    return true;
  }
  void release() { if (p) Dev::free(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

template <class T> bool upload(DevBuf& b, const std::vector<T>& v) {
  if (!b.reserve(v.size() * sizeof(T))) return false;
  return Dev::h2d(b.p, v.data(), v.size() * sizeof(T));
}

struct TimingEntry { const char* name; float ms; int launches; };

}  // namespace

struct relem_batch {
  int nseq = 0;
  int Lmax = 0;
  long long total_len = 0;
  std::vector<long long> off;
  std::vector<unsigned char> kind;
  std::vector<int> gate;
  bool has_gate = false;
  DevBuf d_seq, d_off, d_ws, d_kind, d_gate, d_order;
  long long cells = 0;
};

struct relem_ctx {
  int dev = 0;
  std::string err;
  bool have_energy = false, have_pattern = false, have_params = false;
  EnergyTables et;
  int max_span = 0, max_iloop = 0, no_ene = 0;
  double min_bpp = 0;
  ProfileHMM hmm;
  FlatHMM flat, nullflat;
  int no_rss = 0, no_prf = 0;
  int n_theta = 0;
  DevBuf d_energy, d_codes, d_hmm, d_null, d_n2s, d_theta, d_null_theta;
  DevBuf d_prof;
  DevBuf d_scratch, d_queue, d_res, d_Z, d_ENo, d_ENx, d_EH, d_eff, d_skip;
  DevBuf d_s1, d_s2, d_s3, d_s4, d_s5, d_s6, d_s7, d_s8, d_s9, d_s10;
  DevEnergy den;
  DevEnergy denl;                      // exponentiated copy of the energy tables (linear-space filter pass)
  lin::LinHost linh;                   // transition lists grouped by parent and by child
  lin::LinHMM dlh;
  lin::LinParams dlp;
  lin::LinState* lin = nullptr;
  relem_batch* staging = nullptr;      // reused by the host-buffer entry points
  DevBuf d_energy_lin, d_lin_ints, d_lin_w, d_flag, d_order2;
  bool have_lin = false;
  DevHMM dh, dnull;
  DevParams dpar, dnullpar;
  std::vector<TimingEntry> timing;
#ifndef RELEM_HOST_EMU
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  void* nccl_lib = nullptr;
  void* nccl_comm = nullptr;
  DevBuf d_coll;
#endif
};

namespace {

int fail(relem_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  else g_create_error = msg;
  return code;
}

#ifndef RELEM_HOST_EMU
#define CUDA_TRY(c, expr)                                                                        \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return fail(c, RELEM_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));          \
  } while (0)
#endif

int code_of(const std::string& s, int len) {
  if ((int)s.size() != len) return -1;
  int c = 0;
  for (char ch : s) {
    int b;
    switch (ch) {
      case 'N': b = 0; break; case 'A': b = 1; break; case 'C': b = 2; break; case 'G': b = 3; break;
      case 'U': b = 4; break; default: return -1;
    }
    c = c * 5 + b;
  }
  return c;
}

const int kMaxHairpin = 10002;

bool upload_energy(relem_ctx* c) {
  const EnergyTables& t = c->et;
  std::vector<double> blob;
  auto put = [&](const double* p, size_t n) { size_t o = blob.size(); blob.insert(blob.end(), p, p + n); return o; };
  int HL = std::min(c->max_span, kMaxHairpin - 2) + 2;
  std::vector<double> hl(HL);
  for (int d = 0; d < HL; ++d) hl[d] = t.hairpin_len(d);
  size_t o_hl = put(hl.data(), HL);
  size_t o_mmh = put(&t.mismatch_h[0][0][0], 175), o_mmi = put(&t.mismatch_i[0][0][0], 175);
  size_t o_mmm = put(&t.mismatch_m[0][0][0], 200), o_1ni = put(&t.mismatch_1ni[0][0][0], 175);
  size_t o_23i = put(&t.mismatch_23i[0][0][0], 175), o_ext = put(&t.mismatch_ext[0][0][0], 200);
  size_t o_st = put(&t.stack[0][0], 49), o_bu = put(t.bulge, 31), o_in = put(t.internal, 31), o_ni = put(t.ninio, 31);
  size_t o_d5 = put(&t.dangle5[0][0], 40), o_d3 = put(&t.dangle3[0][0], 40);
  size_t o_11 = put(&t.int11[0][0][0][0], 1600), o_21 = put(&t.int21[0][0][0][0][0], 8000);
  size_t o_22 = put(&t.int22[0][0][0][0][0][0], 40000);
  std::vector<int> codes;
  std::vector<double> w3, w4, w6;
  std::vector<int> c3, c4, c6;
  // entries keep their list position (first match wins); malformed entries can never match a window
  for (size_t k = 0; k < t.tri.size(); ++k) { c3.push_back(code_of(t.tri[k], 5)); w3.push_back(t.tri_w[k]); }
  for (size_t k = 0; k < t.tetra.size(); ++k) { c4.push_back(code_of(t.tetra[k], 6)); w4.push_back(t.tetra_w[k]); }
  for (size_t k = 0; k < t.hexa.size(); ++k) { c6.push_back(code_of(t.hexa[k], 8)); w6.push_back(t.hexa_w[k]); }
  double dummy = 0;
  size_t o_w3 = put(w3.empty() ? &dummy : w3.data(), std::max<size_t>(1, w3.size()));
  size_t o_w4 = put(w4.empty() ? &dummy : w4.data(), std::max<size_t>(1, w4.size()));
  size_t o_w6 = put(w6.empty() ? &dummy : w6.data(), std::max<size_t>(1, w6.size()));
  size_t oc3 = codes.size(); codes.insert(codes.end(), c3.begin(), c3.end()); codes.push_back(-2);
  size_t oc4 = codes.size(); codes.insert(codes.end(), c4.begin(), c4.end()); codes.push_back(-2);
  size_t oc6 = codes.size(); codes.insert(codes.end(), c6.begin(), c6.end()); codes.push_back(-2);
  if (!upload(c->d_energy, blob) || !upload(c->d_codes, codes)) return false;
  const double* b = c->d_energy.as<double>();
  const int* ic = c->d_codes.as<int>();
  DevEnergy& e = c->den;
  e.hairpin_len = b + o_hl; e.mismatch_h = b + o_mmh; e.mismatch_i = b + o_mmi; e.mismatch_m = b + o_mmm;
  e.mismatch_1ni = b + o_1ni; e.mismatch_23i = b + o_23i; e.mismatch_ext = b + o_ext; e.stack = b + o_st;
  e.bulge = b + o_bu; e.internal = b + o_in; e.ninio = b + o_ni; e.dangle5 = b + o_d5; e.dangle3 = b + o_d3;
  e.int11 = b + o_11; e.int21 = b + o_21; e.int22 = b + o_22;
  e.term_au = t.term_au; e.mlintern = t.mlintern; e.mlclosing = t.mlclosing;
  e.tri_code = ic + oc3; e.tri_w = b + o_w3; e.ntri = (int)c3.size();
  e.tetra_code = ic + oc4; e.tetra_w = b + o_w4; e.ntetra = (int)c4.size();
  e.hexa_code = ic + oc6; e.hexa_w = b + o_w6; e.nhexa = (int)c6.size();
  e.no_ene = c->no_ene; e.max_span = c->max_span; e.max_iloop = c->max_iloop;
  e.filter = c->min_bpp != 0. ? 1 : 0;
  e.min_lnbpp = std::log(c->min_bpp);
  // the same tables as Boltzmann factors (exp of the log weights; forbidden = 0)
  std::vector<double> lblob(blob.size());
  for (size_t k = 0; k < blob.size(); ++k) lblob[k] = std::exp(blob[k]);
  if (!upload(c->d_energy_lin, lblob)) return false;
  DevEnergy& l = c->denl;
  l = e;
  const double* lb = c->d_energy_lin.as<double>();
  l.hairpin_len = lb + o_hl; l.mismatch_h = lb + o_mmh; l.mismatch_i = lb + o_mmi; l.mismatch_m = lb + o_mmm;
  l.mismatch_1ni = lb + o_1ni; l.mismatch_23i = lb + o_23i; l.mismatch_ext = lb + o_ext; l.stack = lb + o_st;
  l.bulge = lb + o_bu; l.internal = lb + o_in; l.ninio = lb + o_ni; l.dangle5 = lb + o_d5; l.dangle3 = lb + o_d3;
  l.int11 = lb + o_11; l.int21 = lb + o_21; l.int22 = lb + o_22;
  l.term_au = std::exp(t.term_au); l.mlintern = std::exp(t.mlintern); l.mlclosing = std::exp(t.mlclosing);
  l.tri_w = lb + o_w3; l.tetra_w = lb + o_w4; l.hexa_w = lb + o_w6;
  return true;
}

bool upload_hmm(const FlatHMM& f, DevBuf& buf, DevHMM& d) {
  std::vector<int> blob;
  auto put = [&](const std::vector<int>& v) { size_t o = blob.size(); blob.insert(blob.end(), v.begin(), v.end()); blob.push_back(0); return o; };
  size_t o1 = put(f.st_l), o2 = put(f.st_r), o3 = put(f.is_loop), o4 = put(f.right_off), o5 = put(f.right_idx),
         o6 = put(f.left_off), o7 = put(f.left_idx), o8 = put(f.pair_off), o9 = put(f.pair_idx), o10 = put(f.quad_off),
         o11 = put(f.quad_s1), o12 = put(f.quad_s2), o13 = put(f.quad_s3), o14 = put(f.split_off),
         o15 = put(f.split_left), o16 = put(f.split_right), o17 = put(f.node), o18 = put(f.theta_id),
         o19 = put(f.theta_off), o20 = put(f.right_tgt), o21 = put(f.left_tgt), o22 = put(f.pair_tgt),
         o23 = put(f.quad_tgt), o24 = put(f.split_tgt);
  if (!upload(buf, blob)) return false;
  const int* b = buf.as<int>();
  d.M = f.M; d.S = f.S;
  d.st_l = b + o1; d.st_r = b + o2; d.is_loop = b + o3; d.right_off = b + o4; d.right_idx = b + o5;
  d.left_off = b + o6; d.left_idx = b + o7; d.pair_off = b + o8; d.pair_idx = b + o9; d.quad_off = b + o10;
  d.quad_s1 = b + o11; d.quad_s2 = b + o12; d.quad_s3 = b + o13; d.split_off = b + o14; d.split_left = b + o15;
  d.split_right = b + o16; d.node = b + o17; d.theta_id = b + o18; d.theta_off = b + o19;
  d.right_tgt = b + o20; d.left_tgt = b + o21; d.pair_tgt = b + o22; d.quad_tgt = b + o23; d.split_tgt = b + o24;
  d.n_right = (int)f.right_idx.size(); d.n_left = (int)f.left_idx.size(); d.n_pair = (int)f.pair_idx.size();
  d.n_quad = (int)f.quad_s1.size(); d.n_split = (int)f.split_left.size();
  d.s00 = f.s00; d.s0M2 = f.s0M2; d.s0M1 = f.s0M1;
  return true;
}

struct Launch {
  relem_ctx* c;
  SlotLayout lay;
  int nslots = 0;
};

// --no-rss (the linear profile HMM, RNAelem::compute_inside / compute_outside, motif_model.hpp:170-219): the
// reference walks the exterior row alone -- O(i,s) <- O(i-1,s1) over the right transitions, no base pair anywhere,
// EnergyModel::set_seq never called (bpp_eff stays 0).  That is the structured model with the band collapsed to
// span 0 and the base-pair filter off, so the same kernels serve it at a few launches per batch.
static inline int eff_span(const relem_ctx* c) { return c->no_rss ? 0 : c->max_span; }
static void apply_mode(relem_ctx* c) {
  const int filt = (!c->no_rss && c->min_bpp != 0.) ? 1 : 0;
  c->den.max_span = eff_span(c); c->den.filter = filt;
  c->denl.max_span = eff_span(c); c->denl.filter = filt;
}

// choose the number of resident CTAs (= scratch slots) and make sure the scratch fits
int plan_launch(relem_ctx* c, const relem_batch* b, int nch, bool coupled, const void* kernel, Launch& L,
                int max_n = 1 << 30, int threads = RELEM_CTA_THREADS, bool kernel_is_viterbi = false) {
  int S = c->flat.S, M = c->flat.M;
  L.c = c;
  const FlatHMM& f = c->flat;
  const size_t nlm = std::max(std::max(std::max(f.right_idx.size(), f.left_idx.size()), std::max(f.pair_idx.size(), f.split_left.size())),
                              f.quad_s1.size());
  const bool vit = kernel_is_viterbi;
  int vit_cap = 0;   // limit on the sequences per Viterbi CTA (0 = what shared memory allows)
  if (const char* e = std::getenv("RELEM_VIT_READS")) vit_cap = std::max(1, std::atoi(e));
  L.lay = make_layout(std::max(1, b->Lmax), eff_span(c), S, M, c->n_theta, nch, coupled, (int)nlm, vit ? -(vit_cap + 1) : 0);
  const int sm_bytes = vit ? L.lay.sm_total_vit : L.lay.sm_total;
  const int R = vit ? L.lay.vit_R : 1;   // slots per CTA
  unsigned long long band = (unsigned long long)NPLANE * (b->Lmax + 1) * (L.lay.Wmax + 1) * (coupled ? S : 1);
  if (band >= (1ull << 31)) return fail(c, RELEM_EINVAL, "band table of one sequence exceeds 2^31 entries");
#ifdef RELEM_HOST_EMU
  (void)kernel;
  L.nslots = 1;
#else
  if (sm_bytes > 227 * 1024) return fail(c, RELEM_EINVAL, "sequence too long: masks do not fit shared memory");
  CUDA_TRY(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_bytes));
  int occ = 0;
  CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, sm_bytes));
  if (occ < 1) return fail(c, RELEM_ECUDA, "kernel cannot be resident (occupancy 0)");
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(c, cudaMemGetInfo(&free_b, &total_b));
  size_t avail = free_b + c->d_scratch.bytes;
  size_t per = L.lay.stride * sizeof(double);
  long long by_mem = (long long)((avail * 0.85) / (double)per);
  // nslots = resident CTAs; each owns R slots
  L.nslots = (int)std::min<long long>(std::min<long long>((std::min(b->nseq, max_n) + R - 1) / R, (long long)c->sm_count * occ), by_mem / R);
  if (const char* e = std::getenv("RELEM_MAX_SLOTS")) L.nslots = std::min(L.nslots, std::max(1, std::atoi(e)));
  if (L.nslots < 1) return fail(c, RELEM_ENOMEM, "not enough device memory for one sequence slot");
#endif
  if (!c->d_scratch.reserve((size_t)L.nslots * R * L.lay.stride * sizeof(double)))
    return fail(c, RELEM_ENOMEM, "scratch allocation failed");
  if (!c->d_queue.reserve(sizeof(int)) || !Dev::zero(c->d_queue.p, sizeof(int)))
    return fail(c, RELEM_ENOMEM, "queue allocation failed");
  return RELEM_OK;
}

void model_views(relem_ctx* c, ModelView& nullm, ModelView& m) {
  m.h = c->dh; m.p = c->dpar; m.en = c->den;
  nullm.h = c->dnull; nullm.p = c->dnullpar; nullm.en = c->den;
}

BatchView batch_view(const relem_batch* b) {
  BatchView v;
  v.nseq = b->nseq; v.seq = b->d_seq.as<unsigned char>(); v.off = b->d_off.as<long long>();
  v.ws = b->d_ws.as<double>(); v.kind = b->d_kind.as<unsigned char>(); v.order = b->d_order.as<int>();
  return v;
}

#ifndef RELEM_HOST_EMU
struct Timer {
  relem_ctx* c; const char* name; cudaEvent_t a, b;
  Timer(relem_ctx* c_, const char* n) : c(c_), name(n) {
    cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, c->stream);
  }
  void stop() {
    cudaEventRecord(b, c->stream); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    c->timing.push_back(TimingEntry{name, ms, 1});
    cudaEventDestroy(a); cudaEventDestroy(b);
  }
};
#else
struct Timer { Timer(relem_ctx* c, const char* n) { c->timing.push_back(TimingEntry{n, 0.f, 1}); } void stop() {} };
#endif

int ready(relem_ctx* c) {
  if (!c) return RELEM_EINVAL;
  if (!c->have_energy) return fail(c, RELEM_EINVAL, "relem_set_energy has not been called");
  if (!c->have_pattern) return fail(c, RELEM_EINVAL, "relem_set_pattern has not been called");
  if (!c->have_params) return fail(c, RELEM_EINVAL, "relem_set_params has not been called");
  return RELEM_OK;
}

}  // namespace

// ======================================================================================================= ABI
extern "C" {

// every entry point may be called from a fresh host thread (one thread per context in a multi-GPU process):
// make the context's GPU current before touching device memory
static inline bool bind_device(const relem_ctx* c) {
#ifndef RELEM_HOST_EMU
  return cudaSetDevice(c->dev) == cudaSuccess;
#else
  (void)c; return true;
#endif
}

const char* relem_version(void) {
#ifdef RELEM_HOST_EMU
  return "relem-b200 0.1 (HOST EMULATION - debug only)";
#else
  return "relem-b200 0.1 (sm_100a)";
#endif
}

int relem_create(relem_ctx** out, int device_ordinal) {
  if (!out) return RELEM_EINVAL;
  *out = nullptr;
#ifndef RELEM_HOST_EMU
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, RELEM_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                          " (librelem has no CPU path)");
  if (device_ordinal < 0 || device_ordinal >= n) return fail(nullptr, RELEM_EINVAL, "bad device ordinal");
  e = cudaSetDevice(device_ordinal);
  if (e != cudaSuccess) return fail(nullptr, RELEM_ECUDA, cudaGetErrorString(e));
#endif
  relem_ctx* c = new relem_ctx();
  c->dev = device_ordinal;
#ifndef RELEM_HOST_EMU
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) { delete c; return fail(nullptr, RELEM_ECUDA, "cudaGetDeviceProperties"); }
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreate(&c->stream) != cudaSuccess) { delete c; return fail(nullptr, RELEM_ECUDA, "cudaStreamCreate"); }
#endif
  c->nullflat.null_model();
  std::vector<double> z4(4, 0.);
  if (!upload_hmm(c->nullflat, c->d_null, c->dnull) || !upload(c->d_null_theta, z4)) {
    delete c;
    return fail(nullptr, RELEM_ENOMEM, "device allocation failed");
  }
  c->dnullpar.theta = c->d_null_theta.as<double>(); c->dnullpar.n_theta = 4;
  c->dnullpar.lambda0 = 1.; c->dnullpar.lambda1 = 1.; c->dnullpar.ltau = 0.; c->dnullpar.no_prf = 1;
  *out = c;
  return RELEM_OK;
}

void relem_destroy(relem_ctx* c) {
  if (!c) return;
#ifndef RELEM_HOST_EMU
  cudaSetDevice(c->dev);
  if (c->nccl_comm && c->nccl_lib) {
    typedef int (*destroy_t)(void*);
    destroy_t f = (destroy_t)dlsym(c->nccl_lib, "ncclCommDestroy");
    if (f) f(c->nccl_comm);
  }
  c->d_coll.release();
#endif
  if (c->lin) lin::lin_state_destroy(c->lin);
  if (c->staging) relem_batch_destroy(c, c->staging);
  DevBuf* all[] = {&c->d_energy_lin, &c->d_lin_ints, &c->d_lin_w, &c->d_flag, &c->d_order2, &c->d_prof, &c->d_energy, &c->d_codes, &c->d_hmm, &c->d_null, &c->d_n2s, &c->d_theta, &c->d_null_theta,
                   &c->d_scratch, &c->d_queue, &c->d_res, &c->d_Z, &c->d_ENo, &c->d_ENx, &c->d_EH, &c->d_eff,
                   &c->d_skip, &c->d_s1, &c->d_s2, &c->d_s3, &c->d_s4, &c->d_s5, &c->d_s6, &c->d_s7, &c->d_s8,
                   &c->d_s9, &c->d_s10};
  for (DevBuf* b : all) b->release();
#ifndef RELEM_HOST_EMU
  if (c->stream) cudaStreamDestroy(c->stream);
#endif
  delete c;
}

const char* relem_last_error(const relem_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int relem_set_energy(relem_ctx* c, const char* param, int max_span, int max_iloop, double min_bpp, int no_ene) {
  if (!c || !param) return RELEM_EINVAL;
  if (!bind_device(c)) return fail(c, RELEM_ECUDA, "cudaSetDevice failed");
  if (max_span < 0 || min_bpp < 0) return fail(c, RELEM_EINVAL, "bad max_span / min_bpp");
  EnergyInts ints;
  std::string name(param);
  if (!builtin_energy_ints(name, ints)) {
    std::ifstream ifs(name.c_str());
    if (!ifs) return fail(c, RELEM_EIO, "cannot open param file: " + name);
    std::stringstream ss; ss << ifs.rdbuf();
    std::string err;
    if (!parse_param_text(ss.str(), ints, err)) return fail(c, RELEM_EIO, "failed parsing energy parameter: " + err);
  }
  c->et.build(ints);
  c->max_span = max_span; c->max_iloop = max_iloop; c->min_bpp = min_bpp; c->no_ene = no_ene ? 1 : 0;
  if (!upload_energy(c)) return fail(c, RELEM_ENOMEM, "energy upload failed");
  c->have_energy = true;
  apply_mode(c);
  return RELEM_OK;
}

int relem_set_pattern(relem_ctx* c, const char* pattern, int no_rss, int no_prf) {
  if (!c || !pattern) return RELEM_EINVAL;
  if (!bind_device(c)) return fail(c, RELEM_ECUDA, "cudaSetDevice failed");
  if (no_rss && no_prf) return fail(c, RELEM_EINVAL, "no-rss, no-profile are exclusive.");
  try {
    c->hmm.build(pattern);
  } catch (std::exception& e) {
    return fail(c, RELEM_EINVAL, e.what());
  }
  if (no_rss && std::string(pattern).find(')') != std::string::npos)
    return fail(c, RELEM_EINVAL, "search pattern must not include pair when no-rss mode");
  c->flat.from(c->hmm);
  c->no_rss = no_rss ? 1 : 0; c->no_prf = no_prf ? 1 : 0;
  if (c->have_energy) apply_mode(c);
  c->n_theta = c->hmm.n_theta();
  std::vector<int> n2s;
  for (auto& r : c->hmm.n2s) n2s.insert(n2s.end(), r.begin(), r.end());
  if (!upload_hmm(c->flat, c->d_hmm, c->dh) || !upload(c->d_n2s, n2s) ||
      !c->d_theta.reserve(sizeof(double) * c->n_theta))
    return fail(c, RELEM_ENOMEM, "automaton upload failed");
  c->linh.build(c->flat);
  if (!upload(c->d_lin_ints, c->linh.blob)) return fail(c, RELEM_ENOMEM, "automaton upload failed");
  c->linh.view(c->d_lin_ints.as<int>(), c->dlh);
  c->have_pattern = true;
  c->have_params = false;
  return RELEM_OK;
}

int relem_model_dims(const relem_ctx* c, int* M, int* S, int* n_rows, int* n_theta) {
  if (!c || !c->have_pattern) return RELEM_EINVAL;
  if (M) *M = c->hmm.M;
  if (S) *S = c->hmm.S;
  if (n_rows) *n_rows = (int)c->hmm.row_size.size();
  if (n_theta) *n_theta = c->n_theta;
  return RELEM_OK;
}

int relem_theta_rows(const relem_ctx* c, int* row_sizes) {
  if (!c || !c->have_pattern || !row_sizes) return RELEM_EINVAL;
  for (size_t k = 0; k < c->hmm.row_size.size(); ++k) row_sizes[k] = c->hmm.row_size[k];
  return RELEM_OK;
}

int relem_hmm_get(const relem_ctx* c, int kind, int* out) {
  if (!c || !c->have_pattern) return -1;
  const ProfileHMM& h = c->hmm;
  std::vector<int> v;
  auto csr = [&](const std::vector<std::vector<int>>& lists) {
    int o = 0;
    v.push_back(0);
    for (auto& r : lists) { o += (int)r.size(); v.push_back(o); }
    for (auto& r : lists) v.insert(v.end(), r.begin(), r.end());
  };
  switch (kind) {
    case 0: for (auto& s : h.state) { v.push_back(s.l); v.push_back(s.r); } break;
    case 1: v = h.loop_state; break;
    case 2: csr(h.right); break;
    case 3: csr(h.left); break;
    case 4: csr(h.pairt); break;
    case 5: for (auto& q : h.quads) v.insert(v.end(), q.begin(), q.end()); break;
    case 6: v = h.node; break;
    case 7: v = h.theta_id; break;
    case 9: for (auto& r : h.reachable) for (char x : r) v.push_back(x); break;
    default: return -1;
  }
  if (out) std::copy(v.begin(), v.end(), out);
  return (int)v.size();
}

int relem_energy_get(const relem_ctx* c, const char* name, double* out, int cap) {
  if (!c || !c->have_energy || !name) return -1;
  const EnergyTables& t = c->et;
  struct Ent { const char* n; const double* p; int len; };
  const Ent ents[] = {
      {"hairpin", t.hairpin, 31}, {"mismatch_h", &t.mismatch_h[0][0][0], 175}, {"mismatch_i", &t.mismatch_i[0][0][0], 175},
      {"mismatch_m", &t.mismatch_m[0][0][0], 175}, {"mismatch_1ni", &t.mismatch_1ni[0][0][0], 175},
      {"mismatch_23i", &t.mismatch_23i[0][0][0], 175}, {"mismatch_ext", &t.mismatch_ext[0][0][0], 175},
      {"stack", &t.stack[0][0], 49}, {"bulge", t.bulge, 31}, {"term_au", &t.term_au, 1},
      {"int11", &t.int11[0][0][0][0], 1600}, {"int21", &t.int21[0][0][0][0][0], 8000},
      {"int22", &t.int22[0][0][0][0][0][0], 40000}, {"internal", t.internal, 31}, {"dangle5", &t.dangle5[0][0], 40},
      {"dangle3", &t.dangle3[0][0], 40}, {"ninio", t.ninio, 31}, {"mlintern", &t.mlintern, 1},
      {"mlclosing", &t.mlclosing, 1}, {"ml_base", &t.ml_base, 1}, {"lxc37", &t.lxc37, 1},
      {"triloop", t.tri_w.data(), (int)t.tri_w.size()}, {"tetraloop", t.tetra_w.data(), (int)t.tetra_w.size()},
      {"hexaloop", t.hexa_w.data(), (int)t.hexa_w.size()}};
  for (const Ent& e : ents)
    if (!std::strcmp(e.n, name)) {
      int n = std::min(cap, e.len);
      if (out) for (int k = 0; k < n; ++k) out[k] = e.p[k];
      return e.len;
    }
  return -1;
}

int relem_set_params(relem_ctx* c, const double* theta_flat, int n_theta, const double lambda[2], double tau) {
  if (!c || !theta_flat || !lambda) return RELEM_EINVAL;
  if (!bind_device(c)) return fail(c, RELEM_ECUDA, "cudaSetDevice failed");
  if (!c->have_pattern) return fail(c, RELEM_EINVAL, "relem_set_pattern has not been called");
  if (n_theta != c->n_theta) return fail(c, RELEM_EINVAL, "theta size does not match the pattern");
  if (!Dev::h2d(c->d_theta.p, theta_flat, sizeof(double) * n_theta)) return fail(c, RELEM_ECUDA, "theta upload failed");
  c->dpar.theta = c->d_theta.as<double>(); c->dpar.n_theta = n_theta;
  c->dpar.lambda0 = lambda[0]; c->dpar.lambda1 = lambda[1]; c->dpar.ltau = std::log(tau); c->dpar.no_prf = c->no_prf;
  {
    // linear-space tables: per-base scale kappa = exp(-mean background log emission), so that background
    // emissions are O(1) factors
    double lk = 0.;
    if (!c->no_prf && !c->hmm.row_size.empty()) {
      double m = 0.;
      int r0 = c->hmm.row_size[0], cnt = 0;
      for (int k = 0; k < r0; ++k) if (std::isfinite(theta_flat[k])) { m += theta_flat[k]; ++cnt; }
      if (cnt) lk = -m / cnt;
    }
    std::vector<double> wb;
    size_t o_r, o_l, o_p;
    lin::build_lin_weights(c->flat, theta_flat, tau, c->no_prf, std::exp(lk), wb, o_r, o_l, o_p);
    if (!upload(c->d_lin_w, wb)) return fail(c, RELEM_ENOMEM, "weight upload failed");
    const double* b = c->d_lin_w.as<double>();
    c->dlp.r_w = b + o_r; c->dlp.l_w = b + o_l; c->dlp.p_w = b + o_p;
    c->dlp.lambda0 = lambda[0]; c->dlp.lambda1 = lambda[1]; c->dlp.kappa = std::exp(lk); c->dlp.ln_kappa = lk;
    c->dlp.no_prf = c->no_prf; c->dlp.n_theta = n_theta;
    c->have_lin = true;
  }
  c->have_params = true;
  return RELEM_OK;
}

static int batch_fill(relem_ctx* c, relem_batch* b, int nseq, const uint8_t* seq_cat, const int64_t* off,
                      const double* ws_cat, const uint8_t* kind, const int32_t* gate) {
  if (!c || !b || nseq < 0 || !off || (nseq > 0 && (!seq_cat || !ws_cat))) return RELEM_EINVAL;
  if (!c->have_energy) return fail(c, RELEM_EINVAL, "relem_set_energy has not been called");
  b->Lmax = 0; b->cells = 0; b->has_gate = false;
  b->nseq = nseq;
  b->off.assign(off, off + nseq + 1);
  b->total_len = off[nseq] - off[0];
  if (off[0] != 0) { return fail(c, RELEM_EINVAL, "off[0] must be 0"); }
  std::vector<int> order(nseq);
  std::iota(order.begin(), order.end(), 0);
  for (int n = 0; n < nseq; ++n) {
    long long L = off[n + 1] - off[n];
    if (L < 1 || L > 9999) { return fail(c, RELEM_EINVAL, "sequence length must be in 1..9999"); }
    b->Lmax = std::max<int>(b->Lmax, (int)L);
    long long W = std::min<long long>(L, c->max_span);
    b->cells += (L + 1) * (W + 1) - W * (W + 1) / 2;
  }
  for (long long k = 0; k < b->total_len; ++k)
    if (seq_cat[k] > 4) { return fail(c, RELEM_EINVAL, "base codes must be 0..4"); }
  std::stable_sort(order.begin(), order.end(),
                   [&](int a, int bb) { return off[a + 1] - off[a] > off[bb + 1] - off[bb]; });
  b->kind.assign(nseq, 0);
  if (kind) b->kind.assign(kind, kind + nseq);
  b->gate.assign(nseq, -1);
  if (gate) { b->gate.assign(gate, gate + nseq); b->has_gate = true; }
  for (int n = 0; n < nseq; ++n) {
    if (b->kind[n] > 4) { return fail(c, RELEM_EINVAL, "bad sequence kind"); }
    if (b->gate[n] >= nseq || b->gate[n] == n) { return fail(c, RELEM_EINVAL, "bad gate index"); }
  }
  // the caller's buffers go to the device as they are (no staging copies)
  std::vector<long long> offv(off, off + nseq + 1);
  const size_t nb = (size_t)b->total_len;
  bool up = b->d_seq.reserve(nb) && Dev::h2d(b->d_seq.p, seq_cat, nb) && b->d_ws.reserve(nb * sizeof(double)) &&
            Dev::h2d(b->d_ws.p, ws_cat, nb * sizeof(double));
  if (!up || !upload(b->d_off, offv) || !upload(b->d_kind, b->kind) ||
      !upload(b->d_gate, b->gate) || !upload(b->d_order, order)) {
    return fail(c, RELEM_ENOMEM, "batch upload failed");
  }
  return RELEM_OK;
}

int relem_batch_create(relem_ctx* c, int nseq, const uint8_t* seq_cat, const int64_t* off, const double* ws_cat,
                       const uint8_t* kind, const int32_t* gate, relem_batch** out) {
  if (!out || !c) return RELEM_EINVAL;
  if (!bind_device(c)) return fail(c, RELEM_ECUDA, "cudaSetDevice failed");
  relem_batch* b = new relem_batch();
  int rc = batch_fill(c, b, nseq, seq_cat, off, ws_cat, kind, gate);
  if (rc) { relem_batch_destroy(c, b); return rc; }
  *out = b;
  return RELEM_OK;
}

void relem_batch_destroy(relem_ctx*, relem_batch* b) {
  if (!b) return;
  b->d_seq.release(); b->d_off.release(); b->d_ws.release(); b->d_kind.release(); b->d_gate.release();
  b->d_order.release();
  delete b;
}

int64_t relem_batch_cells(const relem_batch* b) { return b ? b->cells : 0; }

int relem_estep_run(relem_ctx* c, relem_batch* b, relem_estep_out* out) {
  int rc = ready(c);
  if (rc) return rc;
  if (!b || !out) return RELEM_EINVAL;
  c->timing.clear();
  const int NT = c->n_theta, nseq = b->nseq;
  out->fn = 0; out->EH_diff[0] = out->EH_diff[1] = 0; out->sum_eff = 0; out->n_skipped = 0;
  if (out->EN_diff) std::fill(out->EN_diff, out->EN_diff + NT, 0.);
  if (nseq == 0) return RELEM_OK;
#ifndef RELEM_HOST_EMU
  CUDA_TRY(c, cudaSetDevice(c->dev));
#endif
  int nres = 7 + 2 * NT;
  if (!c->d_Z.reserve(sizeof(double) * 3 * nseq) || !c->d_ENo.reserve(sizeof(double) * (size_t)NT * nseq) ||
      !c->d_ENx.reserve(sizeof(double) * (size_t)NT * nseq) || !c->d_EH.reserve(sizeof(double) * 4 * nseq) ||
      !c->d_eff.reserve(sizeof(double) * nseq) || !c->d_skip.reserve(nseq) || !c->d_res.reserve(sizeof(double) * nres))
    return fail(c, RELEM_ENOMEM, "output allocation failed");
  EstepOut eo;
  eo.Z = c->d_Z.as<double>(); eo.ENo = c->d_ENo.as<double>(); eo.ENx = c->d_ENx.as<double>();
  eo.EH = c->d_EH.as<double>(); eo.bpp_eff = c->d_eff.as<double>(); eo.skipped = c->d_skip.as<unsigned char>();
  eo.prof = nullptr;
  if (c->d_prof.reserve(16 * sizeof(unsigned long long)) && Dev::zero(c->d_prof.p, 16 * sizeof(unsigned long long)))
    eo.prof = c->d_prof.as<unsigned long long>();
  ModelView nullm, m;
  model_views(c, nullm, m);
  BatchView bv = batch_view(b);
  // Throughput path: scaled linear-space kernels (relem_lin.cu).  Sequences whose scaled values leave the fp64
  // range come back flagged and are re-run below on the log-space kernel, which also serves RELEM_PATH=log.
  const char* path_env = std::getenv("RELEM_PATH");
  bool use_lin = c->have_lin && !(path_env && std::strcmp(path_env, "log") == 0);
  int n_fallback = nseq;
  if (use_lin) {
    if (!c->lin) c->lin = lin::lin_state_create();
    if (!c->d_flag.reserve(nseq) || !Dev::zero(c->d_flag.p, nseq)) return fail(c, RELEM_ENOMEM, "flag allocation failed");
    lin::LinLaunch ll;
    ll.h = c->dlh; ll.p = c->dlp; ll.en = c->den; ll.el = c->denl;
    ll.kappa0 = std::exp(-0.3);
    ll.b = bv; ll.Lmax = b->Lmax; ll.max_span = eff_span(c); ll.out = eo;
    ll.flag = c->d_flag.as<unsigned char>();
    ll.nch = (out->ENo || out->ENx || out->EH) ? 2 : 1;
    ll.max_slots = 0;
    if (const char* e = std::getenv("RELEM_MAX_SLOTS")) ll.max_slots = std::max(1, std::atoi(e));
#ifndef RELEM_HOST_EMU
    ll.sm_count = c->sm_count; ll.stream = (void*)c->stream;
#else
    ll.sm_count = 1; ll.stream = nullptr;
#endif
    float ms = 0.f; int nl = 0; std::string lerr;
    int lrc = lin::lin_estep_launch(c->lin, ll, &ms, &nl, lerr);
    if (lrc == 1) {
      // the automaton's lists do not fit the linear-space kernels' shared memory: whole batch on the log-space path
      use_lin = false;
    } else {
      if (lrc) return fail(c, lrc == 3 ? RELEM_ENOMEM : RELEM_ECUDA, "linear-space E-step: " + lerr);
      c->timing.push_back(TimingEntry{"relem_estep_lin_kernel", ms, nl});
      {
        // RELEM_PHASE_TIMING=1: per-phase-class device time, reported with zero launches so that sums over entries
        // with launches > 0 stay the whole-call figures
        const char* pn[16]; float pms[16]; int pl[16];
        int np = lin::lin_phase_timing(c->lin, pn, pms, pl, 16);
        for (int k = 0; k < np; ++k) c->timing.push_back(TimingEntry{pn[k], pms[k], -pl[k]});
      }
      std::vector<unsigned char> flags(nseq);
      if (!Dev::d2h(flags.data(), c->d_flag.p, nseq)) return fail(c, RELEM_ECUDA, "flag copy failed");
      std::vector<int> redo;
      for (int k = 0; k < nseq; ++k) if (flags[k]) redo.push_back(k);
      n_fallback = (int)redo.size();
      if (n_fallback) {
        if (!upload(c->d_order2, redo)) return fail(c, RELEM_ENOMEM, "fallback list upload failed");
        bv.order = c->d_order2.as<int>();
        bv.nseq = n_fallback;
      }
    }
  }
  if (n_fallback > 0) {
    Launch L;
#ifdef RELEM_HOST_EMU
    rc = plan_launch(c, b, 2, true, nullptr, L, n_fallback);
#else
    rc = plan_launch(c, b, 2, true, (const void*)relem_estep_kernel, L, n_fallback);
#endif
    if (rc) return rc;
    Timer t(c, "relem_estep_kernel");
#ifdef RELEM_HOST_EMU
    std::vector<unsigned char> smem(L.lay.sm_total + 64);
    relem_estep_kernel(nullm, m, bv, L.lay, c->d_scratch.as<double>(), c->d_queue.as<int>(), eo, smem.data());
#else
    relem_estep_kernel<<<L.nslots, RELEM_CTA_THREADS, L.lay.sm_total, c->stream>>>(
        nullm, m, bv, L.lay, c->d_scratch.as<double>(), c->d_queue.as<int>(), eo);
    CUDA_TRY(c, cudaGetLastError());
#endif
    t.stop();
  }
  bv = batch_view(b);
  // the reference never runs the base-pair filter in --no-rss mode: EnergyModel::bpp_eff() keeps its initial 0
  if (c->no_rss && !Dev::zero(c->d_eff.p, sizeof(double) * nseq)) return fail(c, RELEM_ECUDA, "bpp_eff reset failed");
  {
    Timer t(c, "relem_reduce_kernel");
    const int* gp = b->has_gate ? b->d_gate.as<int>() : nullptr;
#ifdef RELEM_HOST_EMU
    relem_gate_kernel(nseq, gp, eo, nullptr);
    for (int k = 0; k < nres; ++k) relem_reduce_kernel(nseq, NT, bv.kind, eo, c->d_res.as<double>(), k, nullptr);
#else
    relem_gate_kernel<<<std::max(1, std::min(1024, (nseq + 127) / 128)), RELEM_CTA_THREADS, 0, c->stream>>>(nseq, gp, eo);
    relem_reduce_kernel<<<nres, RELEM_CTA_THREADS, 0, c->stream>>>(nseq, NT, bv.kind, eo, c->d_res.as<double>(), 0);
    CUDA_TRY(c, cudaGetLastError());
#endif
    t.stop();
  }
#ifndef RELEM_HOST_EMU
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
#endif
  std::vector<double> res(nres);
  if (!Dev::d2h(res.data(), c->d_res.p, sizeof(double) * nres)) return fail(c, RELEM_ECUDA, "result copy failed");
  if (eo.prof) {
    unsigned long long pr[16];
    if (Dev::d2h(pr, eo.prof, sizeof(pr))) {
      static const char* pn[9] = {"phase:prepare+bpp_filter", "phase:emit+inside", "phase:zero_Q", "phase:outside", "phase:fold+store",
                                  "phase:k0_inside", "phase:k0_ext", "phase:k0_outside", "phase:ext_rows"};
      double tot = 0;
      for (int k = 0; k < 9; ++k) tot += (double)pr[k];
      for (int k = 0; k < 9; ++k) c->timing.push_back(TimingEntry{pn[k], tot > 0 ? (float)(pr[k] / tot) : 0.f, 0});
    }
  }
  out->fn = res[0]; out->sum_eff = res[1]; out->n_skipped = (int64_t)res[2];
  out->EH_diff[0] = res[3] - res[5]; out->EH_diff[1] = res[4] - res[6];
  if (out->EN_diff) for (int t = 0; t < NT; ++t) out->EN_diff[t] = res[7 + t] - res[7 + NT + t];
  bool ok = true;
  if (out->Z) ok = ok && Dev::d2h(out->Z, c->d_Z.p, sizeof(double) * 3 * nseq);
  if (out->ENo) ok = ok && Dev::d2h(out->ENo, c->d_ENo.p, sizeof(double) * (size_t)NT * nseq);
  if (out->ENx) ok = ok && Dev::d2h(out->ENx, c->d_ENx.p, sizeof(double) * (size_t)NT * nseq);
  if (out->EH) ok = ok && Dev::d2h(out->EH, c->d_EH.p, sizeof(double) * 4 * nseq);
  if (out->bpp_eff) ok = ok && Dev::d2h(out->bpp_eff, c->d_eff.p, sizeof(double) * nseq);
  if (out->skipped) ok = ok && Dev::d2h(out->skipped, c->d_skip.p, nseq);
  if (!ok) return fail(c, RELEM_ECUDA, "detail copy failed");
  return RELEM_OK;
}

int relem_estep(relem_ctx* c, int nseq, const uint8_t* seq_cat, const int64_t* off, const double* ws_cat,
                const uint8_t* kind, const int32_t* gate, relem_estep_out* out) {
  if (!c) return RELEM_EINVAL;
  if (!bind_device(c)) return fail(c, RELEM_ECUDA, "cudaSetDevice failed");
  // the staging batch lives in the context: its device buffers only grow, so a training loop that calls this
  // once per iteration pays for the copies, not for allocations
  if (!c->staging) c->staging = new relem_batch();
  int rc = batch_fill(c, c->staging, nseq, seq_cat, off, ws_cat, kind, gate);
  if (rc) return rc;
  return relem_estep_run(c, c->staging, out);
}

int relem_bpp(relem_ctx* c, relem_batch* b, int64_t* moff, uint8_t* bp_ok, uint8_t* left_ok, double* lnbpp,
              double* bpp_eff, double* lnZ) {
  if (!c || !b || !moff) return RELEM_EINVAL;
  if (!c->have_energy) return fail(c, RELEM_EINVAL, "relem_set_energy has not been called");
  if (c->no_rss) return fail(c, RELEM_EINVAL, "no base-pair filter in --no-rss mode");
  c->timing.clear();
  const int nseq = b->nseq;
  std::vector<long long> mo(nseq + 1, 0);
  for (int n = 0; n < nseq; ++n) {
    long long L = b->off[n + 1] - b->off[n], W = std::min<long long>(L, c->max_span);
    mo[n + 1] = mo[n] + (L + 1) * (W + 1);
  }
  for (int n = 0; n <= nseq; ++n) moff[n] = mo[n];
  if (nseq == 0) return RELEM_OK;
#ifndef RELEM_HOST_EMU
  CUDA_TRY(c, cudaSetDevice(c->dev));
#endif
  // the filter pass needs no motif: run it with the null automaton in both model slots
  ModelView nullm, m;
  nullm.h = c->dnull; nullm.p = c->dnullpar; nullm.en = c->den;
  m = nullm;
  FlatHMM keep = c->flat; int keep_nt = c->n_theta;
  c->flat = c->nullflat; c->n_theta = 4;
  Launch L;
#ifdef RELEM_HOST_EMU
  int rc = plan_launch(c, b, 1, false, nullptr, L);
#else
  int rc = plan_launch(c, b, 1, false, (const void*)relem_bpp_kernel, L);
#endif
  c->flat = keep; c->n_theta = keep_nt;
  if (rc) return rc;
  size_t tot = (size_t)mo[nseq];
  if (!c->d_s1.reserve(sizeof(long long) * (nseq + 1)) || !c->d_s2.reserve(tot) || !c->d_s3.reserve(tot) ||
      !c->d_s4.reserve(lnbpp ? tot * sizeof(double) : 8) || !c->d_s5.reserve(sizeof(double) * nseq) ||
      !c->d_s6.reserve(sizeof(double) * nseq))
    return fail(c, RELEM_ENOMEM, "output allocation failed");
  Dev::h2d(c->d_s1.p, mo.data(), sizeof(long long) * (nseq + 1));
  BppOut bo;
  bo.moff = c->d_s1.as<long long>(); bo.bp_ok = c->d_s2.as<unsigned char>(); bo.left_ok = c->d_s3.as<unsigned char>();
  bo.lnbpp = lnbpp ? c->d_s4.as<double>() : nullptr; bo.bpp_eff = c->d_s5.as<double>(); bo.lnZ = c->d_s6.as<double>();
  Dev::zero(c->d_s6.p, sizeof(double) * nseq);
  BatchView bv = batch_view(b);
  {
    Timer t(c, "relem_bpp_kernel");
#ifdef RELEM_HOST_EMU
    std::vector<unsigned char> smem(L.lay.sm_total + 64);
    relem_bpp_kernel(nullm, m, bv, L.lay, c->d_scratch.as<double>(), c->d_queue.as<int>(), bo, smem.data());
#else
    relem_bpp_kernel<<<L.nslots, RELEM_CTA_THREADS, L.lay.sm_total, c->stream>>>(nullm, m, bv, L.lay, c->d_scratch.as<double>(),
                                                                    c->d_queue.as<int>(), bo);
    CUDA_TRY(c, cudaGetLastError());
#endif
    t.stop();
  }
#ifndef RELEM_HOST_EMU
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
#endif
  bool ok = true;
  if (bp_ok) ok = ok && Dev::d2h(bp_ok, bo.bp_ok, tot);
  if (left_ok) ok = ok && Dev::d2h(left_ok, bo.left_ok, tot);
  if (lnbpp) ok = ok && Dev::d2h(lnbpp, bo.lnbpp, tot * sizeof(double));
  if (bpp_eff) ok = ok && Dev::d2h(bpp_eff, bo.bpp_eff, sizeof(double) * nseq);
  if (lnZ) ok = ok && Dev::d2h(lnZ, bo.lnZ, sizeof(double) * nseq);
  if (!ok) return fail(c, RELEM_ECUDA, "result copy failed");
  return RELEM_OK;
}

int relem_scan_run(relem_ctx* c, relem_batch* b, relem_scan_out* out) {
  int rc = ready(c);
  if (rc) return rc;
  if (!b || !out) return RELEM_EINVAL;
  c->timing.clear();
  const int NT = c->n_theta, nseq = b->nseq;
  if (out->EN) std::fill(out->EN, out->EN + NT, 0.);
  if (nseq == 0) return RELEM_OK;
#ifndef RELEM_HOST_EMU
  CUDA_TRY(c, cudaSetDevice(c->dev));
#endif
  // slots (= resident CTAs) of the Viterbi kernel; the all-in-one log-space kernel plans its own further down, only
  // when a sequence actually falls back to it
  Launch L;
#ifdef RELEM_HOST_EMU
  rc = plan_launch(c, b, 1, true, nullptr, L, 1 << 30, RELEM_VIT_THREADS, true);
#else
  rc = plan_launch(c, b, 1, true, (const void*)relem_viterbi_kernel, L, 1 << 30, RELEM_VIT_THREADS, true);
#endif
  if (rc) return rc;
  size_t tl = (size_t)b->total_len;
  if (!c->d_s1.reserve(sizeof(double) * tl) || !c->d_s2.reserve(sizeof(double) * (tl + nseq)) ||
      !c->d_s3.reserve(sizeof(double) * tl) || !c->d_s4.reserve(sizeof(int) * tl) || !c->d_s5.reserve(tl) ||
      !c->d_s6.reserve(sizeof(int) * nseq) || !c->d_s7.reserve(sizeof(int) * nseq) ||
      !c->d_s8.reserve(sizeof(double) * nseq) || !c->d_s9.reserve(sizeof(double) * (size_t)NT * nseq) ||
      !c->d_s10.reserve(sizeof(double) * nseq))
    return fail(c, RELEM_ENOMEM, "output allocation failed");
  ScanOut so;
  so.PysL = c->d_s1.as<double>(); so.PyeL = c->d_s2.as<double>(); so.PyiL = c->d_s3.as<double>();
  so.psihat = c->d_s4.as<int>(); so.rss = c->d_s5.as<char>(); so.Ys = c->d_s6.as<int>(); so.Ye = c->d_s7.as<int>();
  so.exist = c->d_s8.as<double>(); so.EN = c->d_s9.as<double>(); so.ZL = c->d_s10.as<double>();
  ModelView nullm, m;
  model_views(c, nullm, m);
  BatchView bv = batch_view(b);
  // Posterior passes on the linear-space kernels, Viterbi (bit-exact) on the log-space kernel; sequences that leave
  // the fp64 range there (flag) and RELEM_PATH=log go through the all-in-one log-space scan kernel.
  const char* path_env = std::getenv("RELEM_PATH");
  bool use_lin = c->have_lin && !(path_env && std::strcmp(path_env, "log") == 0);
  int n_fallback = nseq;
  if (use_lin) {
    if (!c->lin) c->lin = lin::lin_state_create();
    if (!c->d_flag.reserve(nseq) || !Dev::zero(c->d_flag.p, nseq)) return fail(c, RELEM_ENOMEM, "flag allocation failed");
    lin::LinScanLaunch ll;
    ll.h = c->dlh; ll.p = c->dlp; ll.en = c->den; ll.el = c->denl;
    ll.kappa0 = std::exp(-0.3);
    ll.b = bv; ll.Lmax = b->Lmax; ll.max_span = eff_span(c); ll.so = so;
    ll.flag = c->d_flag.as<unsigned char>();
    ll.max_slots = 0;
    if (const char* e = std::getenv("RELEM_MAX_SLOTS")) ll.max_slots = std::max(1, std::atoi(e));
#ifndef RELEM_HOST_EMU
    ll.stream = (void*)c->stream;
    CUDA_TRY(c, cudaFuncSetAttribute((const void*)relem_viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     L.lay.sm_total_vit));
#else
    ll.stream = nullptr;
#endif
    struct VitCtx {
      relem_ctx* c; ModelView m; BatchView bv; Launch* L; ScanOut so;
#ifndef RELEM_HOST_EMU
      std::vector<cudaEvent_t> ev;   // one pair per chunk around the Viterbi launch
#endif
    } vc{c, m, bv, &L, so};
    auto after_chunk = [](void* user, const lin::LinChunkView& cv) -> int {
      VitCtx& v = *(VitCtx*)user;
      ExtMasks em;
      em.scratch = cv.scratch; em.stride = cv.stride; em.masks_off = cv.masks_off; em.mask_words = cv.mask_words;
      em.base = cv.base; em.count = cv.count;
      relem_ctx* c = v.c;
#ifdef RELEM_HOST_EMU
      *c->d_queue.as<int>() = 0;
      std::vector<unsigned char> smem(std::max(v.L->lay.sm_total, v.L->lay.sm_total_vit) + 64);
      relem_viterbi_kernel(v.m, v.bv, v.L->lay, c->d_scratch.as<double>(), c->d_queue.as<int>(), c->d_n2s.as<int>(), v.so,
                           em, c->d_flag.as<unsigned char>(), smem.data());
      return 0;
#else
      cudaStream_t st = (cudaStream_t)cv.stream;
      // chunks alternate between two streams; the Viterbi launches share one value table and one work counter, so each
      // waits for the one before it
      if (!v.ev.empty()) cudaStreamWaitEvent(st, v.ev.back(), 0);
      if (cudaMemsetAsync(c->d_queue.p, 0, sizeof(int), st) != cudaSuccess) return 1;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      v.ev.push_back(e0); v.ev.push_back(e1);
      cudaEventRecord(e0, st);
      relem_viterbi_kernel<<<std::min(v.L->nslots, (cv.count + v.L->lay.vit_R - 1) / v.L->lay.vit_R), RELEM_VIT_THREADS,
                             v.L->lay.sm_total_vit, st>>>(
          v.m, v.bv, v.L->lay, c->d_scratch.as<double>(), c->d_queue.as<int>(), c->d_n2s.as<int>(), v.so, em,
          c->d_flag.as<unsigned char>());
      cudaEventRecord(e1, st);
      return cudaGetLastError() == cudaSuccess ? 0 : 1;
#endif
    };
    float ms = 0.f; int nl = 0; std::string lerr;
    int lrc = lin::lin_scan_launch(c->lin, ll, after_chunk, &vc, &ms, &nl, lerr);
    float vit_ms = 0.f;
    int vit_n = 0;
#ifndef RELEM_HOST_EMU
    if (lrc == 0)
      for (size_t k = 0; k + 1 < vc.ev.size(); k += 2) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, vc.ev[k], vc.ev[k + 1]) == cudaSuccess) { vit_ms += t; ++vit_n; }
      }
    for (cudaEvent_t e : vc.ev) cudaEventDestroy(e);
#endif
    if (lrc == 1) {
      use_lin = false;
    } else {
      if (lrc) return fail(c, lrc == 3 ? RELEM_ENOMEM : RELEM_ECUDA, "linear-space scan: " + lerr);
      // the chunk loop's bracket covers the posterior passes AND the Viterbi launches; the second entry is the
      // Viterbi share of it (0 launches: already counted in the first)
      c->timing.push_back(TimingEntry{"relem_scan_lin_kernels", ms, nl + vit_n});
      c->timing.push_back(TimingEntry{"relem_viterbi_kernel (share of the above)", vit_ms, -vit_n});
      std::vector<unsigned char> flags(nseq);
      if (!Dev::d2h(flags.data(), c->d_flag.p, nseq)) return fail(c, RELEM_ECUDA, "flag copy failed");
      std::vector<int> redo;
      for (int k = 0; k < nseq; ++k) if (flags[k]) redo.push_back(k);
      n_fallback = (int)redo.size();
      if (n_fallback) {
        if (!upload(c->d_order2, redo)) return fail(c, RELEM_ENOMEM, "fallback list upload failed");
        bv.order = c->d_order2.as<int>();
        bv.nseq = n_fallback;
      }
    }
  }
  if (n_fallback > 0) {
    // the all-in-one kernel has its own (full) slot layout
#ifndef RELEM_HOST_EMU
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // the scratch may be re-allocated below
    rc = plan_launch(c, b, 1, true, (const void*)relem_scan_kernel, L, n_fallback);
#else
    rc = plan_launch(c, b, 1, true, nullptr, L, n_fallback);
#endif
    if (rc) return rc;
    if (!Dev::zero(c->d_queue.p, sizeof(int))) return fail(c, RELEM_ECUDA, "queue reset failed");
    Timer t(c, "relem_scan_kernel");
#ifdef RELEM_HOST_EMU
    std::vector<unsigned char> smem(L.lay.sm_total + 64);
    relem_scan_kernel(nullm, m, bv, L.lay, c->d_scratch.as<double>(), c->d_queue.as<int>(), c->d_n2s.as<int>(), so,
                      smem.data());
#else
    relem_scan_kernel<<<std::min(L.nslots, n_fallback), RELEM_CTA_THREADS, L.lay.sm_total, c->stream>>>(
        nullm, m, bv, L.lay, c->d_scratch.as<double>(), c->d_queue.as<int>(), c->d_n2s.as<int>(), so);
    CUDA_TRY(c, cudaGetLastError());
#endif
    t.stop();
  }
#ifndef RELEM_HOST_EMU
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
#endif
  bool ok = true;
  if (out->PysL) ok = ok && Dev::d2h(out->PysL, so.PysL, sizeof(double) * tl);
  if (out->PyeL) ok = ok && Dev::d2h(out->PyeL, so.PyeL, sizeof(double) * (tl + nseq));
  if (out->PyiL) ok = ok && Dev::d2h(out->PyiL, so.PyiL, sizeof(double) * tl);
  if (out->psihat) ok = ok && Dev::d2h(out->psihat, so.psihat, sizeof(int) * tl);
  if (out->rss) ok = ok && Dev::d2h(out->rss, so.rss, tl);
  if (out->Ys) ok = ok && Dev::d2h(out->Ys, so.Ys, sizeof(int) * nseq);
  if (out->Ye) ok = ok && Dev::d2h(out->Ye, so.Ye, sizeof(int) * nseq);
  if (out->exist_prob) ok = ok && Dev::d2h(out->exist_prob, so.exist, sizeof(double) * nseq);
  if (out->ZL) ok = ok && Dev::d2h(out->ZL, so.ZL, sizeof(double) * nseq);
  if (out->EN) {
    // E[N] summed over the batch in input order (RNAelemScanDP::operator(), motif_scanner.hpp:253-258)
    std::vector<double> en((size_t)NT * nseq);
    ok = ok && Dev::d2h(en.data(), so.EN, sizeof(double) * en.size());
    for (int n = 0; n < nseq; ++n) for (int t = 0; t < NT; ++t) out->EN[t] += en[(size_t)n * NT + t];
  }
  if (!ok) return fail(c, RELEM_ECUDA, "result copy failed");
  return RELEM_OK;
}

int relem_scan(relem_ctx* c, int nseq, const uint8_t* seq_cat, const int64_t* off, const double* ws_cat,
               relem_scan_out* out) {
  if (!c) return RELEM_EINVAL;
  if (!bind_device(c)) return fail(c, RELEM_ECUDA, "cudaSetDevice failed");
  // same staging batch as relem_estep: device buffers that only grow.  Allocating and freeing six buffers per call next
  // to a ~100 GB scratch allocation cost 0.15-0.5 s per call (tools/scan_e2e_probe.py).
  if (!c->staging) c->staging = new relem_batch();
  int rc = batch_fill(c, c->staging, nseq, seq_cat, off, ws_cat, nullptr, nullptr);
  if (rc) return rc;
  return relem_scan_run(c, c->staging, out);
}

void relem_assigned_range(int64_t total, int n, int k, int64_t* from, int64_t* to) {
  // contiguous blocks, the remainder spread over the first ranks (arrayjob_manager.hpp:141-149)
  int64_t base = n > 0 ? total / n : 0, rem = n > 0 ? total % n : 0;
  int64_t f = base * k + std::min<int64_t>(k, rem);
  int64_t t = f + base + (k < rem ? 1 : 0);
  if (from) *from = f;
  if (to) *to = t;
}

int relem_last_timing(const relem_ctx* c, const char** names, float* ms, int* launches, int cap) {
  if (!c) return 0;
  int n = std::min<int>(cap, (int)c->timing.size());
  for (int k = 0; k < n; ++k) {
    if (names) names[k] = c->timing[k].name;
    if (ms) ms[k] = c->timing[k].ms;
    if (launches) launches[k] = c->timing[k].launches;
  }
  return n;
}

int relem_fp64_peak(relem_ctx* c, double* dfma_per_s, double* exp_per_s) {
  if (!c) return RELEM_EINVAL;
  if (!bind_device(c)) return fail(c, RELEM_ECUDA, "cudaSetDevice failed");
  std::string err;
#ifndef RELEM_HOST_EMU
  int rc = lin::lin_fp64_peak((void*)c->stream, c->sm_count, dfma_per_s, exp_per_s, err);
#else
  int rc = lin::lin_fp64_peak(nullptr, 1, dfma_per_s, exp_per_s, err);
#endif
  if (rc) return fail(c, rc == 3 ? RELEM_ENOMEM : rc == 1 ? RELEM_EINVAL : RELEM_ECUDA, "fp64 micro-benchmark: " + err);
  return RELEM_OK;
}

// ------------------------------------------------------------------------------------------------- NCCL
#ifdef RELEM_HOST_EMU
int relem_comm_unique_id(uint8_t*) { return RELEM_EINVAL; }
int relem_comm_init(relem_ctx* c, const uint8_t*, int, int) { return fail(c, RELEM_EINVAL, "no collectives in the emulation"); }
int relem_allreduce_sum(relem_ctx* c, double*, int) { return fail(c, RELEM_EINVAL, "no collectives in the emulation"); }
#else
namespace {
void* nccl_open() {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) return h;
  }
  return nullptr;
}
}  // namespace

int relem_comm_unique_id(uint8_t id[128]) {
  void* lib = nccl_open();
  if (!lib) { g_create_error = "libnccl.so.2 not found"; return RELEM_EINVAL; }
  typedef int (*fn_t)(void*);
  fn_t f = (fn_t)dlsym(lib, "ncclGetUniqueId");
  if (!f) { g_create_error = "ncclGetUniqueId not found"; return RELEM_EINVAL; }
  return f(id) == 0 ? RELEM_OK : RELEM_ECUDA;
}

int relem_comm_init(relem_ctx* c, const uint8_t id[128], int rank, int nranks) {
  if (!c || !id) return RELEM_EINVAL;
  CUDA_TRY(c, cudaSetDevice(c->dev));
  if (!c->nccl_lib) c->nccl_lib = nccl_open();
  if (!c->nccl_lib) return fail(c, RELEM_EINVAL, "libnccl.so.2 not found");
  struct Uid { char b[128]; } uid;
  std::memcpy(uid.b, id, 128);
  typedef int (*init_t)(void**, int, Uid, int);
  init_t f = (init_t)dlsym(c->nccl_lib, "ncclCommInitRank");
  if (!f) return fail(c, RELEM_EINVAL, "ncclCommInitRank not found");
  int r = f(&c->nccl_comm, nranks, uid, rank);
  if (r != 0) return fail(c, RELEM_ECUDA, "ncclCommInitRank failed: " + std::to_string(r));
  return RELEM_OK;
}

int relem_allreduce_sum(relem_ctx* c, double* host_buf, int n) {
  if (!c || !host_buf || n < 0) return RELEM_EINVAL;
  if (!c->nccl_comm) return fail(c, RELEM_EINVAL, "relem_comm_init has not been called");
  CUDA_TRY(c, cudaSetDevice(c->dev));
  if (!c->d_coll.reserve(sizeof(double) * n)) return fail(c, RELEM_ENOMEM, "collective buffer");
  CUDA_TRY(c, cudaMemcpyAsync(c->d_coll.p, host_buf, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  typedef int (*ar_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  ar_t f = (ar_t)dlsym(c->nccl_lib, "ncclAllReduce");
  if (!f) return fail(c, RELEM_EINVAL, "ncclAllReduce not found");
  const int ncclDouble = 8, ncclSum = 0;
  int r = f(c->d_coll.p, c->d_coll.p, (size_t)n, ncclDouble, ncclSum, c->nccl_comm, c->stream);
  if (r != 0) return fail(c, RELEM_ECUDA, "ncclAllReduce failed: " + std::to_string(r));
  CUDA_TRY(c, cudaMemcpyAsync(host_buf, c->d_coll.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return RELEM_OK;
}
#endif

}  // extern "C"
