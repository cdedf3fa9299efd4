"""The literal CUDA schedules (ring / ring_chunked / halving_doubling / bcube) are step
tables built on the host; simulate them with numpy (steps are barrier-separated, all
reads of a step see the state left by the previous step) and check the allreduce."""
import numpy as np
import pytest

import gloo_b200 as gb

build = gb._C.cuda.build_schedule


def simulate(name, size, count, base=2, pack=4):
    tables = [build(name, r, size, count, base, pack) for r in range(size)]
    nsteps = {len(t) for t in tables}
    assert len(nsteps) == 1, f"ranks disagree on step count: {nsteps}"
    bufs = [np.arange(count, dtype=np.float64) * size + r for r in range(size)]
    stage = [np.zeros(count) for _ in range(size)]
    for s in range(nsteps.pop()):
        prev = [b.copy() for b in bufs]
        prev_stage = [b.copy() for b in stage]
        writes = [None] * size
        for r in range(size):
            st = tables[r][s]
            lo, hi = st["off"], st["off"] + st["len"]
            assert hi <= count
            if st["len"] == 0:
                continue
            src = prev_stage if st["from_stage"] else prev
            if st["mode"] == 2:    # STAGE
                stage[r][lo:hi] = prev[r][lo:hi]
            elif st["mode"] == 1:  # COPY
                bufs[r][lo:hi] = src[st["peers"][0]][lo:hi]
            else:                  # REDUCE
                acc = prev[r][lo:hi].copy()
                for p in st["peers"]:
                    acc += src[p][lo:hi]
                bufs[r][lo:hi] = acc
            writes[r] = (lo, hi, st["mode"])
        # race check: nobody may read a range that its owner writes in the same step
        for r in range(size):
            st = tables[r][s]
            if st["len"] == 0 or st["mode"] == 2:
                continue
            lo, hi = st["off"], st["off"] + st["len"]
            for p in st["peers"]:
                w = writes[p]
                if w is None or st["from_stage"] != (w[2] == 2):
                    continue
                assert hi <= w[0] or lo >= w[1], f"{name}: step {s}: rank {r} reads [{lo},{hi}) of rank {p} which writes [{w[0]},{w[1]})"
    exp = np.arange(count, dtype=np.float64) * size * size + size * (size - 1) / 2
    for r in range(size):
        np.testing.assert_array_equal(bufs[r], exp, err_msg=f"{name} P={size} n={count} rank {r}")


@pytest.mark.parametrize("name", ["ring", "ring_chunked", "halving_doubling", "bcube"])
@pytest.mark.parametrize("size", [2, 3, 4, 5, 6, 7, 8, 12, 16])
def test_schedule_is_an_allreduce(name, size):
    for count in (1, 7, 64, 1000, 4099):
        simulate(name, size, count)


@pytest.mark.parametrize("base,size", [(3, 9), (4, 16), (4, 8), (3, 6)])
def test_bcube_bases(base, size):
    simulate("bcube", size, 5000, base=base)
    # fewer, wider steps than base 2
    assert len(build("bcube", 0, size, 5000, base, 4)) <= len(build("bcube", 0, size, 5000, 2, 4))


def test_step_counts_match_cost_model():
    # docs/algorithms.md: ring P-1 rounds (+1 publish step), chunked ring 2(P-1), HD 2 lg P
    assert len(build("ring", 0, 8, 1 << 20, 2, 4)) == 8
    assert len(build("ring_chunked", 0, 8, 1 << 20, 2, 4)) == 14
    assert len(build("halving_doubling", 0, 8, 1 << 20, 2, 4)) == 6
    # P = 6 = 4 + 2 (binary blocks): 2 halving steps, 1 chain-up, 1 chain-down phase of 2 pieces, 2 doubling steps
    six = build("halving_doubling", 0, 6, 1 << 20, 2, 4)
    assert len(six) == 7 and sum(s["sync"] for s in six) == 6


try:
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st
except ImportError:  # pragma: no cover - hypothesis is optional
    given = None

if given is not None:

    @settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(name=st.sampled_from(["ring", "ring_chunked", "halving_doubling", "bcube"]),
           size=st.integers(2, 16), count=st.integers(1, 3000), pack=st.sampled_from([1, 2, 4, 8, 16]))
    def test_any_schedule_any_shape(name, size, count, pack):
        """Random (algorithm, rank count, length, vector width): the table is race-free and
        computes the allreduce, including lengths below the rank count and widths that do not
        divide the length."""
        simulate(name, size, count, pack=pack)

    @settings(max_examples=60, deadline=None)
    @given(base=st.integers(2, 5), size=st.integers(2, 16), count=st.integers(1, 2000))
    def test_bcube_any_base(base, size, count):
        simulate("bcube", size, count, base=base)


# ---- pipelined halving-doubling: several steps share one barrier phase (sync flag) ------------

def simulate_phased(name, size, count, chunks=2, pack=4):
    tables = [build(name, r, size, count, chunks, pack) for r in range(size)]
    phases = []
    for t in tables:
        ph = []
        for st in t:
            if st["sync"] or not ph:
                ph.append([])
            ph[-1].append(st)
        phases.append(ph)
    assert len({len(p) for p in phases}) == 1, "ranks disagree on the number of barrier phases"
    bufs = [np.arange(count, dtype=np.float64) * size + r for r in range(size)]
    for k in range(len(phases[0])):
        prev = [b.copy() for b in bufs]
        writes = [[] for _ in range(size)]
        for r in range(size):
            for st in phases[r][k]:
                lo, hi = st["off"], st["off"] + st["len"]
                if st["len"] == 0:
                    continue
                assert hi <= count and not st["from_stage"] and st["mode"] in (0, 1)
                if st["mode"] == 1:
                    bufs[r][lo:hi] = prev[st["peers"][0]][lo:hi]
                else:
                    acc = prev[r][lo:hi].copy()
                    for p in st["peers"]:
                        acc += prev[p][lo:hi]
                    bufs[r][lo:hi] = acc
                for w in writes[r]:
                    assert hi <= w[0] or lo >= w[1], "two steps of one phase write overlapping ranges"
                writes[r].append((lo, hi))
        for r in range(size):
            for st in phases[r][k]:
                lo, hi = st["off"], st["off"] + st["len"]
                for p in st["peers"] if st["len"] else []:
                    for w in writes[p]:
                        assert hi <= w[0] or lo >= w[1], f"phase {k}: rank {r} reads what rank {p} writes"
    exp = np.arange(count, dtype=np.float64) * size * size + size * (size - 1) / 2
    for r in range(size):
        np.testing.assert_array_equal(bufs[r], exp)
    return len(phases[0])


@pytest.mark.parametrize("size", [2, 3, 4, 5, 6, 8, 12, 16])
@pytest.mark.parametrize("chunks", [1, 2, 3, 4])
def test_halving_doubling_pipelined_is_an_allreduce(size, chunks):
    for count in (1, 63, 1000, 40000):
        simulate_phased("halving_doubling_pipelined", size, count, chunks)


def test_halving_doubling_pipelined_overlaps_chunks():
    # 2 chunks skewed by one step: 2 lg P + 1 phases instead of 2 x 2 lg P sequential steps
    plain = len(build("halving_doubling", 0, 8, 1 << 20, 2, 4))
    assert simulate_phased("halving_doubling_pipelined", 8, 1 << 16, 2) == plain + 1
    steps = build("halving_doubling_pipelined", 0, 8, 1 << 16, 2, 4)
    assert len(steps) == 2 * plain and sum(s["sync"] for s in steps) == plain + 1


# ---- halving-doubling on a rank count that is not a power of two: binary blocks vs folding --------

@pytest.mark.parametrize("size", [3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15])
def test_halving_doubling_binary_blocks(size, monkeypatch):
    for count in (1, 5, 64, 1001, 70000):
        phases = simulate_phased("halving_doubling", size, count)
    tables = [build("halving_doubling", r, size, 70000, 2, 4) for r in range(size)]
    blocks = [1 << i for i in range(5, -1, -1) if size & (1 << i)]
    lg = blocks[0].bit_length() - 1
    assert phases == 2 * lg + 2 * (len(blocks) - 1)
    # nobody idles through the halving phase except ranks of blocks that have fewer steps, and the
    # extra traffic of a small block is its own share: no step moves the whole vector unless a
    # block of one rank has to end up with all of it
    if blocks[-1] > 1:
        assert max(st["len"] for t in tables for st in t) <= 70000 // 2 + 4
    # the folding variant (GLB_HD_FOLD=1) is still there and moves the whole vector twice
    monkeypatch.setenv("GLB_HD_FOLD", "1")
    simulate("halving_doubling", size, 1001)
    folded = build("halving_doubling", size - 1, size, 70000, 2, 4)
    assert max(st["len"] for st in folded) == 70000
