# A/B runs of full library variants built by scripts/build_variant_full.sh (experiment helper)
for v in "" b256 b512; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  for w in svf svf biquad; do
    timeout 120 python bench.py --workload $w --mix 0 --no-cpu --steps 50 --warmup 5 2>&1 | python -c "import sys,json; L=sys.stdin.read().splitlines(); J=[l for l in L if l.startswith(chr(123))]; print('variant=${v:-base} $w', (lambda d:(d['value'], round(d['roofline']['frac'],4), d['roofline']['launch_ms_median']))(json.loads(J[-1])) if J else L[-3:])"
  done
done
