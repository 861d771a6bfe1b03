#!/bin/bash
# Dev aid: time the headline step under several settings of the launch knobs (environment variables read by csrc/).
# usage: tools/knob_sweep.sh "<label>|<env assignments>|<extra bench flags>" ...   -> gpurun_out/knob_sweep.txt
mkdir -p gpurun_out
out=gpurun_out/knob_sweep.txt
: > $out
for spec in "$@"; do
  IFS='|' read -r label envs flags <<< "$spec"
  line=$(env $envs timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline $flags 2>gpurun_out/knob_$label.err | tail -1)
  ms=$(python -c "import json,sys; d=json.loads(sys.argv[1]); print(d.get('ms_per_step'), d.get('value'))" "$line" 2>/dev/null)
  echo "$label | $envs | $flags | $ms" | tee -a $out
done
