#!/bin/sh
# Sanitizer pass over the host transport and collectives (the reference's -DSANITIZE=<x>):
# P ranks as processes x 2 threads each, running the instrumented benchmark binary built by
# `python build.py --sanitize <thread|address|undefined>`.
# usage: [GLB_SAN_TRANSPORT=tcp|uv] scripts/sanitizer_run.sh <thread|address|undefined> [P] [benchmark ...]
SAN=${1:-thread}; P=${2:-3}; shift; shift
BIN=gloo_b200/bin/glb_benchmark_$SAN
[ -x $BIN ] || python build.py --sanitize $SAN || exit 1
[ $# -gt 0 ] || set -- allreduce_ring allreduce_ring_chunked allreduce_halving_doubling allreduce_bcube \
  new_allreduce new_allreduce_ring new_allreduce_bcube reduce_scatter new_reduce_scatter allgather allgather_ring \
  alltoall alltoall_v gather scatter reduce broadcast broadcast_one_to_all barrier_all_to_all \
  barrier_all_to_one pairwise_exchange sendrecv_roundtrip sendrecv_stress
OUT=$(mktemp -d /tmp/glb_san.XXXXXX)
for NAME in "$@"; do
  for EL in 1000 500000; do
    D=$(mktemp -d /tmp/glb_san_rdv.XXXXXX)
    r=0; while [ $r -lt $P ]; do
      TSAN_OPTIONS="suppressions=$PWD/.tsan-suppressions halt_on_error=0 second_deadlock_stack=1 log_path=$OUT/$NAME.$EL.r$r" \
      ASAN_OPTIONS="detect_leaks=1 log_path=$OUT/$NAME.$EL.r$r" \
      UBSAN_OPTIONS="print_stacktrace=1 log_path=$OUT/$NAME.$EL.r$r" \
        $BIN --size $P --rank $r --shared-path $D --transport ${GLB_SAN_TRANSPORT:-tcp} --iteration-count 30 --warmup-iters 2 \
             --threads 2 --elements $EL $NAME >$OUT/$NAME.$EL.r$r.stdout 2>&1 &
      r=$((r+1)); done
    wait; rm -rf $D
  done
done
# TLS pairs (session lifetime vs. the loop thread): short runs, many connects and teardowns
if command -v openssl >/dev/null 2>&1; then
  C=$OUT/certs; mkdir -p $C
  openssl req -x509 -newkey rsa:2048 -nodes -keyout $C/ca.key -out $C/ca.crt -subj /CN=san-ca -days 2 >/dev/null 2>&1
  openssl req -newkey rsa:2048 -nodes -keyout $C/r.key -out $C/r.csr -subj /CN=rank >/dev/null 2>&1
  openssl x509 -req -in $C/r.csr -CA $C/ca.crt -CAkey $C/ca.key -CAcreateserial -out $C/r.crt -days 2 >/dev/null 2>&1
  i=0; while [ $i -lt 15 ]; do
    D=$(mktemp -d /tmp/glb_san_rdv.XXXXXX)
    for r in 1 0; do
      TSAN_OPTIONS="suppressions=$PWD/.tsan-suppressions halt_on_error=0 log_path=$OUT/tls.$i.r$r" \
      ASAN_OPTIONS="detect_leaks=0 log_path=$OUT/tls.$i.r$r" \
        $BIN --size 2 --rank $r --shared-path $D --transport tls --pkey $C/r.key --cert $C/r.crt --ca-file $C/ca.crt \
             --elements 1000 --iteration-count 10 new_allreduce_ring >$OUT/tls.$i.r$r.stdout 2>&1 &
    done
    wait; rm -rf $D; i=$((i+1))
  done
fi
# the self-test: same instrumentation, results checked (threads as ranks in one process)
ST=gloo_b200/bin/glb_selftest_$SAN
if [ -x $ST ]; then
  TSAN_OPTIONS="suppressions=$PWD/.tsan-suppressions halt_on_error=0 log_path=$OUT/selftest.r0" \
  ASAN_OPTIONS="detect_leaks=1 log_path=$OUT/selftest.r0" \
  UBSAN_OPTIONS="print_stacktrace=1 log_path=$OUT/selftest.r0" \
    $ST 2 3 4 >$OUT/selftest.stdout 2>&1 || echo "selftest failed under $SAN (see $OUT/selftest.stdout)"
  tail -1 $OUT/selftest.stdout
fi
n=$(cat $OUT/*.r[0-9]*.[0-9]* 2>/dev/null | grep -c "WARNING: ThreadSanitizer\|ERROR: AddressSanitizer\|ERROR: LeakSanitizer\|runtime error")
echo "$SAN sanitizer reports: $n (logs in $OUT)"
[ "$n" = 0 ]
