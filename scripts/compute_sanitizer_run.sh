#!/bin/bash
# compute-sanitizer over the single-process self-tests (every collective kernel once on virtual
# loop-back ranks + the local reduce / broadcast kernels). memcheck catches out-of-bounds and
# misaligned accesses, racecheck the shared-memory hazards of the TMA ring, synccheck barrier misuse.
# One process, one kernel at a time: kernels of different ranks that wait for each other cannot run
# under the sanitizer (it serialises launches), which is exactly what the loop-back ranks avoid.
# usage: scripts/compute_sanitizer_run.sh [outdir]   (needs a GPU)
out=${1:-gpurun_out}
mkdir -p "$out"
cat > /tmp/glb_sanitize_target.py <<'PY'
import sys
import torch
import gloo_b200 as gb
from gloo_b200.ops import cuda as gcu


def single(ctx):
    cc = gcu.CudaContext(ctx, 0, stage_bytes=64 << 20)
    stream = torch.cuda.current_stream().cuda_stream
    res = cc.pc.loopback_selftest(stream, 1 << 15)
    bad = [r for r in res if not r["ok"]]
    for r in res:
        print(("skip" if r["skipped"] else "ok" if r["ok"] else "FAIL"), r["name"], r["detail"], flush=True)
    assert not bad, bad
    return len(res)


print("self-tests run:", gb.spawn_threads(1, single, cuda_device=0))
res = gb._C.cuda.local_ops_selftest([0], 20011)
assert all(r["ok"] for r in res), res
print("local ops:", len(res), "ok")
PY
rc=0
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --report-api-errors no --print-limit 20 \
    env PYTHONPATH=$PWD python /tmp/glb_sanitize_target.py > "$out/sanitizer_$tool.log" 2>&1
  r=$?
  echo "compute-sanitizer $tool rc=$r" | tee -a "$out/sanitizer_summary.txt"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" "$out/sanitizer_$tool.log" | tail -3 | tee -a "$out/sanitizer_summary.txt"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
