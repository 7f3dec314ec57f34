# run bench.py once per tuning variant under variants/ (differt_b200/build.py --variant ...)
for v in variants/*/; do n=$(basename $v); DIFFERT_B200_LIB=$PWD/$v/libdiffert_b200.so timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n', 'value %.4g'%d['value'], 'executed %.4g'%d['executed_tests_per_s'], 'frac %.3f'%d['executed_fraction_of_algorithmic'], 'step %.1f ms'%d['ms_per_step'], 'blockage %.1f ms'%d['roofline']['kernel_ms'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"; done
