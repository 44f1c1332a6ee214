# diagnostics: end-to-end GCUPS against the pipelining chunk size of gmapdp_run_batch
for mb in 64 128 192 384 768; do
  echo "== GMAPDP_CHUNK_MB=$mb"
  GMAPDP_CHUNK_MB=$mb timeout 300 python bench.py --no-cpu-baseline --chain-problems 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('resident %.1f ms  e2e %.1f ms  (%.1f GCUPS)' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value']))"
done
