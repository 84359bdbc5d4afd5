for f1 in 1.08 1.18 1.3; do for f in 1.12 1.18 1.25; do
python scripts/run_config.py C horizon_factor_first=$f1 horizon_factor=$f --reps 2 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['opts'], 'loop %.1f total %.1f raises %d n_exact %d refine %.1f digest %s' % (d['ms_loop'], d['ms_total'], d['n_horizon_raises'], d['n_exact'], d['ms_refine'], d['digest'][:8]))"
done; done
