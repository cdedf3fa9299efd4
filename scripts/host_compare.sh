#!/bin/sh
# Config 1 of BASELINE.json on the CPU: host allreduce over TCP loopback, world_size=2,
# file rendezvous — reference `benchmark` vs `glb_benchmark`, same flags.
# usage: scripts/host_compare.sh <benchmark name> [size]
NAME=${1:-allreduce_ring}; P=${2:-2}
REF=baseline/_ref/bin/benchmark; OURS=gloo_b200/bin/glb_benchmark
for BIN in "$REF" "$OURS"; do
  D=$(mktemp -d /tmp/hostcmp.XXXXXX)
  echo "== $BIN $NAME (P=$P)"
  r=1; while [ $r -lt $P ]; do
    $BIN --size $P --rank $r --shared-path $D --transport tcp --iteration-time 500ms $NAME >/dev/null 2>&1 &
    r=$((r+1)); done
  $BIN --size $P --rank 0 --shared-path $D --transport tcp --iteration-time 500ms $NAME 2>&1 | grep -E "^ +[0-9]"
  wait; rm -rf $D
done
