#!/bin/bash
# A/B of library builds on the headline bench: tools/ab_libs.sh libA.so libB.so ...  (paths relative to inmost-fem_b200/)
for rep in 1; do
for l in "$@"; do
  r=$(AFB_LIB=$PWD/inmost-fem_b200/$l python bench.py --steps 10 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f ms/step  gather %.3f  geom %.3f' % (d['ms_per_step'], d['roofline']['step']['gather_ms'], d['roofline']['step']['element_ms']))")
  echo "$l -> $r"
done
done
