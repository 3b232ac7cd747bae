for blk in 1 16 64 256 4096; do
rm -rf /tmp/dump; mkdir -p /tmp/dump
PMC_RENUMBER_BLOCK=$blk PMC_DUMP_SELL=/tmp/dump/op python - <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from common import hex_problem, make_context
c = make_context(hex_problem(16, 3), True, 1e-6, 1e-12, 300)
c.close()
PY
echo "block $blk"
for f in $(ls /tmp/dump/*r17152_c17152_w.bin | head -1) $(ls /tmp/dump/*r17152_c17152_p.bin | head -1); do ./tools/micro/saddle.exe $f | grep "448 threads, staged.*296"; done
PMC_RENUMBER_BLOCK=$blk python tools/ab_stage.py - | cut -c42-110
done
