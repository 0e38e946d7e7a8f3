"""The tile decoder's algorithm, restated in numpy (tests/tile_model.py), against the oracle: hand-built LZ4
blocks and oracle-compressed data, two output alignments, and the two extreme schedules of the per-byte source
step (every earlier row visible / nothing visible) -- the decoded bytes must not depend on the schedule."""
import numpy as np
import pytest

import kat
import synth_streams as ss
import tile_model as tm


def _blocks():
    for profile in ss.PROFILES:
        yield f"synth-{profile}", ss.lz4_block(profile, 120_000, 7)


@pytest.mark.parametrize("name,stream", list(_blocks()))
@pytest.mark.parametrize("schedule", ["front", "none"])
def test_model_decodes_synthetic_blocks(oracle, name, stream, schedule):
    want = oracle.decompress(stream, kat.LZ4, 400_000)
    assert want is not None
    for a in (0, 5):
        st = {}
        got = tm.decode_lz4_block(stream, len(want), a=a, schedule=schedule, stats=st)
        assert got == want, (name, schedule, a)
        assert st["groups"] > 0


@pytest.mark.parametrize("gen", ["text", "mixed", "period7"])
def test_model_decodes_oracle_streams(oracle, gen):
    data = kat.make_input(gen, 200_000)                      # below the RAP threshold: one frame-less block
    stream = oracle.compress(data, kat.LZ4)
    front, none = {}, {}
    assert tm.decode_lz4_block(stream, len(data), a=3, schedule="front", stats=front) == data.tobytes()
    assert tm.decode_lz4_block(stream, len(data), a=3, schedule="none", stats=none) == data.tobytes()
    assert front["rounds"] <= none["rounds"]                 # looking through written entries never costs rounds


@pytest.mark.parametrize("profile", list(ss.PROFILES))
@pytest.mark.parametrize("schedule", ["16", "none"])
def test_model_decodes_synthetic_snappy_streams(oracle, profile, schedule):
    stream = ss.snappy_stream(profile, 120_000, 9)
    want = oracle.decompress(stream, kat.SNAPPY, 400_000)
    assert want is not None
    for a in (0, 11):
        assert tm.decode_snappy_stream(stream, a=a, schedule=schedule) == want, (profile, schedule, a)


@pytest.mark.parametrize("gen", ["log", "text"])
def test_model_decodes_oracle_snappy_streams(oracle, gen):
    data = kat.make_input(gen, 200_000)                      # below the RAP threshold: one raw stream
    stream = oracle.compress(data, kat.SNAPPY)
    st = {}
    assert tm.decode_snappy_stream(stream, a=9, schedule="16", stats=st) == data.tobytes()
    assert st["groups"] > 0
