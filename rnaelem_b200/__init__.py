"""rnaelem_b200 -- thin Python (ctypes) binding of librelem.so, the sm_100a implementation of RNAelem's
inside / outside / expected-count / posterior / Viterbi hot path.

The product is the CUDA library behind the C ABI of include/relem.h; this package only loads it and offers
the small amount of host glue the tests and bench.py need (FASTQ / train.model readers that follow
RNAelem/fastq_io.hpp:64-108 and RNAelem/motif_io.hpp:118-262, and RNAelem::set_ws, motif_model.hpp:62-70).
There is no CPU implementation here: importing works anywhere, but creating a context without a usable
GPU raises RelemError.
"""
from .binding import (RelemError, Context, load_library, lib_path, EstepResult, ScanResult,  # noqa: F401
                      POS_WITHOUT, POS_WITH, NEG, LR_WITHOUT, LR_NEG)
from .hostio import (read_fastq, read_model, seq_codes, quality_to_ws, model_theta_flat,  # noqa: F401
                     band_cells, pack_batch)

__all__ = ["RelemError", "Context", "load_library", "lib_path", "EstepResult", "ScanResult", "read_fastq",
           "read_model", "seq_codes", "quality_to_ws", "model_theta_flat", "band_cells", "pack_batch",
           "POS_WITHOUT", "POS_WITH", "NEG", "LR_WITHOUT", "LR_NEG"]
