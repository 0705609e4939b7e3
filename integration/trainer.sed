# integration/trainer.sed -- the reference-side change to motif_trainer.hpp, as an edit script for a build-time COPY of
# the header (oracle/Makefile target `gpu`; nothing of the reference is stored in this repository).
# 1. pull the binding in once the TR_* mode bits and ushuffle are declared
s/^  class RNAelemTrainDP {$/}\
#include "relem_host_train.hpp"\
namespace iyak {\
  class RNAelemTrainDP {/
# 2. RNAelemTrainer::operator(): the thread fan-out  ClassThread<RNAelemTrainDP> ct(...); ct(fn,gr);  (:617-621)
/^        ClassThread<RNAelemTrainDP>$/,/^        ct(fn,gr);$/c\
        relem_host::estep(*_motif,_qr,_mode,_cnt,_kmer_shuf,_from,_to,_sum_eff,fn,gr);
