for v in variants/*/; do n=$(basename $v); DIFFERT_B200_LIB=$PWD/$v/libdiffert_b200.so timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n', '%.4g'%d['value'], '%.1f ms'%d['roofline']['kernel_ms'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"; done
