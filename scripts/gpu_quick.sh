#!/bin/bash
# quick GPU check: parity on small cases (fused path, no debug snapshots) + bench value
export DCB_VERBOSE=1
DIAG_NODEBUG=1 timeout 300 python scripts/gpu_diag.py small variants 2>&1 | grep -E "=== case|forward ok|logits|bases match|Error|error|dcb200"
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench: value %.0f e2e %.0f ffn %.1f TF/s avg %.3f ms launches/step %d' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['avg_launch_ms'], d['gpu_launches']/d['steps'])); print('   ms/step:', d['roofline']['kernel_ms_per_step'])
    elif 'dcb200' in l or 'rror' in l: print(l.strip())
"
