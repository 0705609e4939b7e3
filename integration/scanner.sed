# integration/scanner.sed -- the reference-side change to motif_scanner.hpp (build-time copy, see trainer.sed)
# 1. pull the binding in
0,/^namespace iyak {$/s//#include "relem_host.hpp"\
namespace iyak {/
# 2. RNAelemScanner::scan: the thread fan-out  ClassThread<RNAelemScanDP> ct(...); ct(_EN);  (:943-946)
/^      ClassThread<RNAelemScanDP> ct(_thread,\*_motif,_mx_input,_mx_output,$/,/^      ct(_EN);$/c\
      relem_host::scan(*_motif,_qr,_out,_EN);
