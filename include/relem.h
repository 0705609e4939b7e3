/* relem.h -- C ABI of librelem.so: the B200 (sm_100a) implementation of RNAelem's per-sequence
 * inside / outside / expected-count / posterior / Viterbi hot path.
 *
 * The reference (iyak/RNAelem) has no FFI; the seam this library replaces is the pair of C++ entry points
 *   int  RNAelemTrainer::operator()(V const& x, double& fn, V& gr)   RNAelem/motif_trainer.hpp:595-633
 *   void RNAelemScanner::scan(RNAelem& model)                        RNAelem/motif_scanner.hpp:938-949
 * whose bodies fan a batch of sequences out to RNAelemTrainDP (motif_trainer.hpp:38-459) and RNAelemScanDP
 * (motif_scanner.hpp:19-914).  Each entry point below names the reference code it stands in for.
 *
 * Conventions: every call returns 0 on success and a non-zero RELEM_E* code otherwise (never throws across
 * the ABI; relem_last_error() gives the message -- the host wrapper turns it into the reference's die(),
 * util.hpp:121-126).  The caller owns every host buffer; the context owns all device memory.  Calls are
 * synchronous.  A context is bound to one GPU and is not thread-safe: use one context per host thread / rank.
 * There is no CPU fallback: relem_create fails when no CUDA device is usable.
 */
#ifndef RELEM_H
#define RELEM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RELEM_OK 0
#define RELEM_EINVAL 1  /* bad argument / call order */
#define RELEM_ECUDA 2   /* CUDA runtime failure      */
#define RELEM_ENOMEM 3  /* device memory exhausted   */
#define RELEM_EIO 4     /* file problem              */

typedef struct relem_ctx relem_ctx;
typedef struct relem_batch relem_batch;

/* sequence kinds: which restricted boundary condition the second outside pass uses and which partition
 * functions must be finite (motif_trainer.hpp:204-245). */
#define RELEM_POS_WITHOUT 0 /* user sequence flagged "does not contain the motif": Zx = Z(ari=0,nasi=1) */
#define RELEM_POS_WITH 1    /* user sequence flagged "contains the motif" (quality string ends in '!')    */
#define RELEM_NEG 2         /* shuffled negative: Zx = Z(0,1); only Z(1,1) must be finite                 */
/* --lik-ratio objective (motif_trainer.hpp:156-202): every sequence contrasts Z(1,1) with Z(1,0) and both must be
 * finite.  A user sequence flagged "contains the motif" is RELEM_POS_WITH (Zo = Z(1,1), Zx = Z(1,0)); one flagged
 * "does not" and every shuffled negative have the roles swapped (Zo = Z(1,0), Zx = Z(1,1)): their fn, EN_diff and
 * EH_diff contributions enter the batch sums with the opposite sign.  The per-sequence detail arrays keep the
 * (1,1)-then-(1,0) order for these kinds. */
#define RELEM_LR_WITHOUT 3  /* user sequence, counts towards sum_eff                                      */
#define RELEM_LR_NEG 4      /* shuffled negative                                                          */

const char* relem_version(void);

/* ---- context ------------------------------------------------------------------------------------------ */
int relem_create(relem_ctx** out, int device_ordinal);
void relem_destroy(relem_ctx*);
const char* relem_last_error(const relem_ctx*); /* ctx may be NULL: error of the last failed relem_create */

/* ---- model -------------------------------------------------------------------------------------------- */
/* RNAelem::set_energy_params (motif_model.hpp:72-78) + EnergyModel::set_param_file (energy_model.hpp:153-161).
 * param: "~T2004~", "~A2007~" or the path of a ViennaRNA-2.0 format parameter file. */
int relem_set_energy(relem_ctx*, const char* param, int max_span, int max_iloop, double min_bpp, int no_ene);

/* RNAelem::set_motif_pattern (motif_model.hpp:80-97): builds the ProfileHMM automaton (profile_hmm.hpp:206-463)
 * on the host and uploads its flattened transition lists. */
int relem_set_pattern(relem_ctx*, const char* pattern, int no_rss, int no_prf);
/* automaton sizes: M nodes, S interval states, n_rows theta rows, n_theta = sum of row sizes */
int relem_model_dims(const relem_ctx*, int* M, int* S, int* n_rows, int* n_theta);
/* row_sizes[n_rows] (4 for background and '.', 6 for ')') */
int relem_theta_rows(const relem_ctx*, int* row_sizes);
/* kind: 0 states (l,r pairs) 1 loop-state ids 2 right (CSR: S+1 offsets then ids) 3 left 4 pair
 *       5 quads (4 ids each, list order) 6 node chars 7 theta_id 9 reachable (M*M).
 * returns the number of ints written (needed when out==NULL), <0 on error */
int relem_hmm_get(const relem_ctx*, int kind, int* out);
/* energy table introspection (log-Boltzmann weights, the reference's array shapes; names as in
 * energy_param.hpp:61-85 without the underscore). returns count copied, <0 if unknown */
int relem_energy_get(const relem_ctx*, const char* name, double* out, int cap);

/* RNAelem::unpack_params (motif_model.hpp:159-168) after the optional softmax: theta_flat = concatenated rows
 * of log emission weights, lambda[2] = {background, inside-motif}, tau as given on the command line. */
int relem_set_params(relem_ctx*, const double* theta_flat, int n_theta, const double lambda[2], double tau);

/* ---- batches ------------------------------------------------------------------------------------------ */
/* Upload a batch: seq_cat = base codes 0..4 (N,A,C,G,U; bio_sequence.hpp:30-41) of all sequences back to
 * back, off[nseq+1] their offsets, ws_cat = log position weights (RNAelem::set_ws, motif_model.hpp:62-70;
 * same offsets, L entries per sequence), kind[nseq] one of RELEM_POS_* / RELEM_NEG, gate[nseq] = index of the
 * sequence whose "skipped" status also drops this one (the positive a negative was shuffled from,
 * motif_trainer.hpp:211-245), or -1.  kind/gate may be NULL (all RELEM_POS_WITHOUT / -1). */
int relem_batch_create(relem_ctx*, int nseq, const uint8_t* seq_cat, const int64_t* off, const double* ws_cat,
                       const uint8_t* kind, const int32_t* gate, relem_batch** out);
void relem_batch_destroy(relem_ctx*, relem_batch*);
int64_t relem_batch_cells(const relem_batch*); /* sum over sequences of (L+1)(W+1) - W(W+1)/2 band cells */

/* ---- E-step (RNAelemTrainDP::operator(), motif_trainer.hpp:124-272, non lik-ratio branch) ------------ */
typedef struct relem_estep_out {
  double fn;          /* sum over kept sequences of Zo - Zx                                   (:226,:244) */
  double* EN_diff;    /* [n_theta] ENo - ENx, theta shaped (before the softmax chain rule)    (:263-265)  */
  double EH_diff[2];  /* EHo - EHx by lambda slot {s.l==s.r, else}; the host merges the slots when
                         lambda[0]==lambda[1] as the reference's value test does              (:380-381)  */
  double sum_eff;     /* sum of bpp_eff over kept user sequences                              (:227)      */
  int64_t n_skipped;  /* sequences dropped for a non-finite partition function                (:211-215)  */
  /* optional per-sequence detail (each may be NULL) */
  double* Z;          /* [nseq][3] Z(1,1), Z(1,0), Z(0,1)                                                 */
  double* ENo;        /* [nseq][n_theta]                                                                  */
  double* ENx;        /* [nseq][n_theta]                                                                  */
  double* EH;         /* [nseq][4] EHo[0],EHo[1],EHx[0],EHx[1]                                            */
  double* bpp_eff;    /* [nseq]                                                                           */
  uint8_t* skipped;   /* [nseq] 1 = own partition function non-finite, 2 = dropped through gate           */
} relem_estep_out;

/* device-resident batch */
int relem_estep_run(relem_ctx*, relem_batch*, relem_estep_out* out);
/* host buffers in, host results out (= batch_create + estep_run + batch_destroy) */
int relem_estep(relem_ctx*, int nseq, const uint8_t* seq_cat, const int64_t* off, const double* ws_cat,
                const uint8_t* kind, const int32_t* gate, relem_estep_out* out);

/* energy-only base-pair filter alone (EnergyModel::set_seq -> fill_bpp_tables, energy_model.hpp:211-276):
 * bp_ok / left_ok are (L+1)*(W+1) bytes per sequence at byte offset moff[n] (moff[nseq+1] is filled by the
 * call: W = min(L,max_span)), index i*(W+1)+d.  lnbpp (same indexing, doubles; may be NULL) receives
 * ln BPP of every canonical pair when min_bpp>0.  Either mask may be NULL. */
int relem_bpp(relem_ctx*, relem_batch*, int64_t* moff, uint8_t* bp_ok, uint8_t* left_ok, double* lnbpp,
              double* bpp_eff, double* lnZ);

/* ---- scan (RNAelemScanDP::operator(), motif_scanner.hpp:215-260) --------------------------------------- */
typedef struct relem_scan_out {
  double* PysL;       /* [sum L]      log start posteriors, sequence n at off[n]             (:204-205)   */
  double* PyeL;       /* [sum (L+1)]  log end posteriors, sequence n at off[n]+n             (:195-202)   */
  double* PyiL;       /* [sum L]      log inner posteriors                                                */
  int32_t* psihat;    /* [sum L]      Viterbi motif node per base                            (:172-184)   */
  char* rss;          /* [sum L]      Viterbi structure chars O L R H I B M, ' ' untouched                */
  int32_t* Ys;        /* [nseq] */
  int32_t* Ye;        /* [nseq] */
  double* exist_prob; /* [nseq] exp(logsum PysL)                                             (:246)       */
  double* EN;         /* [n_theta] expected emission counts summed over the batch            (:253-258)   */
  double* ZL;         /* [nseq] optional (NULL ok): ln Z(1,1)                                             */
} relem_scan_out;
int relem_scan_run(relem_ctx*, relem_batch*, relem_scan_out* out);
int relem_scan(relem_ctx*, int nseq, const uint8_t* seq_cat, const int64_t* off, const double* ws_cat,
               relem_scan_out* out);

/* ---- multi-GPU: the one collective of the path ------------------------------------------------------- */
/* The reference sums fn / gr / sum_eff of its array-job slaves through text files
 * (motif_array_trainer.hpp:20-61).  Here each rank owns one context; relem_comm_init takes the NCCL unique id
 * (128 bytes, generated by rank 0 with relem_comm_unique_id and distributed by the caller, e.g. through
 * torch.distributed) and relem_allreduce_sum adds n doubles in place over NVLink. */
int relem_comm_unique_id(uint8_t id[128]);
int relem_comm_init(relem_ctx*, const uint8_t id[128], int rank, int nranks);
int relem_allreduce_sum(relem_ctx*, double* host_buf, int n);
/* contiguous sharding of `total` items over n ranks, remainder to the first ranks
 * (ArrayJobManager::assigned_range, arrayjob_manager.hpp:141-149) */
void relem_assigned_range(int64_t total, int n, int k, int64_t* from, int64_t* to);

/* ---- instrumentation ---------------------------------------------------------------------------------- */
/* milliseconds the kernels of the last *_run call took (CUDA events on the launch stream) and the number
 * of kernel launches it made; names[i] are static strings.  returns the number of entries (<= cap). */
int relem_last_timing(const relem_ctx*, const char** names, float* ms, int* launches, int cap);

/* fp64 pipe peaks of this GPU by micro-benchmark (a few ms; SURVEY.md 8d asks for them because MEASURED_PEAKS.json
 * only has HBM and bf16): dfma_per_s = double-precision fused multiply-adds per second with every SM busy
 * (8 independent chains per thread), exp_per_s = libm-accurate exp() evaluations per second.  The reference has no
 * counterpart: its inner loops spend their time in logsumexp (util.hpp:195-202), two libm calls per term. */
int relem_fp64_peak(relem_ctx*, double* dfma_per_s, double* exp_per_s);

#ifdef __cplusplus
}
#endif
#endif
