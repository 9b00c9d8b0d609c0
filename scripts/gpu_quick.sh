# quick GPU iteration: parity tests + short device-timed bench lines; BLZ_AB="opt1 opt2" runs one bench per option string
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/quick_pytest.log
for opt in ${BLZ_AB:-default}; do
  echo "== $opt"
  [ "$opt" = default ] && opt=""
  BLZ_OPTIONS=$opt timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -3 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('value %.3e ms/step %.4f frac %.3f kernel_ms %s early_draws %s clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['detail']['kernel_ms'], d['detail']['early_draws'], d['clocks']))
"
done
