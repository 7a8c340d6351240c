"""The oracle (oracle/liboracle.so, our CPU restatement) against the golden vectors in tests/golden/,
which were produced by the unmodified reference C (tools/make_golden.py).  Runs anywhere."""
import glob
import hashlib
import json
import os

import pytest

import helpers

GOLDEN = os.path.join(helpers.ROOT, "tests", "golden")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.json")))


def load_case(name):
    with open(os.path.join(GOLDEN, name + ".es"), "rb") as f:
        es = f.read()
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        info = json.load(f)
    assert hashlib.sha256(es).hexdigest() == info["es_sha256"]
    return es, info


def check_against_golden(lib, name, chunked=False):
    es, info = load_case(name)
    packets = [(0.0, es)] if not chunked else [(i / 30.0, es[o:o + 4096]) for i, o in enumerate(range(0, len(es), 4096))]
    frames, idx, d = helpers.decode_all(lib, packets)
    assert (d.width, d.height, d.codedSize) == (info["width"], info["height"], info["coded_size"])
    assert round(float(d.frameRate), 3) == info["frame_rate"]
    assert len(frames) == len(info["pictures"])
    for k, ((y, cr, cb), g) in enumerate(zip(frames, info["pictures"])):
        assert idx[k] == g["index"], (name, k)
        got = [hashlib.sha256(p.tobytes()).hexdigest() for p in (y, cr, cb)]
        assert got == [g["y"], g["cr"], g["cb"]], f"{name}: picture {k} differs from the reference"
    d.destroy()


def test_golden_present():
    assert len(CASES) >= 10


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    check_against_golden(helpers.oracle_lib(), name)


@pytest.mark.parametrize("name", ["rows_ip", "ffmpeg_176x144_ip"])
def test_oracle_golden_chunked_writes(name):
    """Writing the stream in many pieces (EXPAND buffer growth) before decoding changes nothing."""
    check_against_golden(helpers.oracle_lib(), name, chunked=True)
