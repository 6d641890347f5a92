#!/bin/bash
# Step kernel sweep: the streaming kernel (groups x ring slots) and variant builds of the plain thread-per-env kernel.
mkdir -p gpurun_out/r2
out=gpurun_out/r2/sweep_stream.log
: > $out
wls=${WLS:-"c2 c3 c2w"}
CS_STREAM=0 timeout 100 python tools/sweep_step.py $wls 2>&1 | grep value | sed 's/^/stream=0 /' >> $out
for v in cooperative-search_b200/csrc/variants/*.so; do
  [ -f "$v" ] || continue
  CS_STREAM=0 COOPSEARCH_LIB=$v timeout 100 python tools/sweep_step.py $wls 2>&1 | grep -i "value\|error" | sed "s/^/stream=0 $(basename $v) /" >> $out
done
CFGS=${CFGS:-6:10 6:8 5:10 4:10 3:10}
for cfg in $CFGS; do
  g=${cfg%%:*}; s=${cfg##*:}
  CS_STREAM_GROUPS=$g CS_STREAM_SLOTS=$s timeout 100 python tools/sweep_step.py $wls 2>&1 | grep -i "value\|error" | sed "s/^/groups=$g slots=$s /" >> $out
done
cat $out
