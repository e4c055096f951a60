#!/bin/bash
# bench.py at several Swin chunk sizes (images per pass through the backbone)
for c in "$@"; do
  python bench.py --steps 10 --warmup 3 --no-cpu --swin-chunk $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('chunk $c', round(d['value'],1), 'cap/s', round(d['ms_per_step'],2), 'ms/step  e2e', round(d['e2e']['value'],1), ' gemm TF/s', round(d['roofline']['achieved'],1))"
done
