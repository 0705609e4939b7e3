for s in 148 296 444 592 888; do RELEM_MAX_SLOTS=$s timeout 200 python bench.py --nseq 1024 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('slots',$s,'seq_evals/s',round(d['seq_evals_per_s'],1), d['phase_share'])"; done
