rm -rf /tmp/dump; mkdir -p /tmp/dump
PMC_DUMP_SELL=/tmp/dump/op python - <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from common import hex_problem, make_context
c = make_context(hex_problem(16, 3), True, 1e-6, 1e-12, 300)
c.close()
PY
./tools/micro/saddle.exe $(ls /tmp/dump/*r17152_c17152_w.bin | head -1) | grep -E "apply only|update only|fused|448 threads, staged.*296"
