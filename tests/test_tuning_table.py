"""The measured tuning table: parsing, range lookup, fall-through (host only)."""
import gloo_b200 as gb

cu = gb._C.cuda

TABLE = """
# comment line
allreduce P=8 buf=sym maxbytes=16384 algo=ll blocks=4
allreduce P=8 buf=sym maxbytes=262144 algo=one_shot blocks=8   # trailing comment
allreduce P=8 buf=sym maxbytes=inf algo=nvls blocks=148 unroll=8
allreduce P=8 buf=user maxbytes=inf algo=pipelined blocks=128 tile=1024
allgather P=8 buf=reg maxbytes=4096 algo=ll blocks=0
allgather P=8 buf=reg maxbytes=inf algo=push blocks=222
this line is garbage
allreduce P=x buf=sym maxbytes=1 algo=ll blocks=1
"""


def test_parse_lookup_and_dump():
    cu.tuning_clear()
    try:
        assert cu.tuning_load_string(TABLE) == 6
        e = cu.tuning_lookup("allreduce", 8, 0, 1000)
        assert e["algo"] == "ll" and e["blocks"] == 4
        assert cu.tuning_lookup("allreduce", 8, 0, 16384)["algo"] == "ll"
        assert cu.tuning_lookup("allreduce", 8, 0, 16385)["algo"] == "one_shot"
        big = cu.tuning_lookup("allreduce", 8, 0, 400_000_000)
        assert big["algo"] == "nvls" and big["blocks"] == 148 and big["unroll"] == 8
        assert cu.tuning_lookup("allreduce", 8, 2, 1 << 20)["tile"] == 1024
        assert cu.tuning_lookup("allreduce", 4, 0, 1000) is None      # no group for P=4
        assert cu.tuning_lookup("allreduce", 8, 1, 1000) is None      # no group for registered buffers
        assert cu.tuning_lookup("allgather", 8, 1, 100)["algo"] == "ll"
        dump = cu.tuning_dump()
        assert "allreduce P=8 buf=sym maxbytes=inf algo=nvls blocks=148 unroll=8" in dump
        cu.tuning_clear()
        assert cu.tuning_load_string(dump) == 6  # round trip
    finally:
        cu.tuning_clear()


def test_packaged_table_parses():
    import os

    path = os.path.join(os.path.dirname(gb.__file__), "tuning", "b200.tune")
    if not os.path.exists(path):
        return
    cu.tuning_clear()
    try:
        assert cu.tuning_load_file(path) > 0
    finally:
        cu.tuning_clear()
