#!/bin/bash
# A/B of environment switches on the headline bench: tools/ab_env.sh "VAR=1" "" ...
for e in "$@"; do
  r=$(env $e python bench.py --steps 10 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f ms/step  gather %.3f  geom %.3f' % (d['ms_per_step'], d['roofline']['step']['gather_ms'], d['roofline']['step']['element_ms']))")
  echo "[$e] -> $r"
done
