run() { env "$@" python bench.py --no-cpu-baseline --no-scan --steps 2 --warmup 1 2>/dev/null | python -c "
import sys,json
d=json.loads([x for x in sys.stdin if x.startswith('{')][-1])
ph=d['roofline']['top_kernel']['phases']
print('$*', round(d['seq_evals_per_s']), {k[23:26].strip('> '):round(v) for k,v in ph.items()})"; }
run A=0
run RELEM_TILE_Q=64
run RELEM_TILE_Q=128
run RELEM_TILE_E=64
run RELEM_TILE_E=128
run RELEM_TILE_P=32
run RELEM_TILE_P=64
run RELEM_TILE_D=32
run RELEM_TILE_D=128
run RELEM_TILE_K0=64
