"""World-size-2 test of the N>1 path on CPU (gloo): streams are sharded round-robin over ranks,
every rank decodes its own shard independently (here with the oracle standing in for the GPU),
and only the counters are reduced (sum of frames, max of seconds).  No data-path collective."""
import hashlib
import json
import os
import subprocess
import sys

import helpers

WORKER = r"""
import hashlib, json, os, sys
sys.path.insert(0, os.environ["REPO"]); sys.path.insert(0, os.path.join(os.environ["REPO"], "tests")); sys.path.insert(0, os.path.join(os.environ["REPO"], "tools"))
import torch.distributed as dist
import helpers, synth_es
from jsmpeg_b200.shard import assign_streams, aggregate
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
names = sorted(synth_es.CASES)[:6]
mine = assign_streams(len(names), rank, world)
digests, frames = {}, 0
for i in mine:
    fr, idx, d = helpers.decode_all(helpers.oracle_lib(), [(0.0, synth_es.make_case(names[i]))])
    frames += len(fr)
    h = hashlib.sha256()
    for planes in fr:
        for p in planes: h.update(p.tobytes())
    digests[names[i]] = h.hexdigest()
    d.destroy()
total, seconds = aggregate(frames, 1.0 + rank)
gathered = [None] * world
dist.all_gather_object(gathered, digests)
if rank == 0:
    merged = {}
    for g in gathered: merged.update(g)
    print("RESULT " + json.dumps({"total": total, "seconds": seconds, "digests": merged}))
dist.destroy_process_group()
"""


def test_two_ranks_partition_the_streams(tmp_path):
    import synth_es
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    import socket
    with socket.socket() as sock:  # a free rendezvous port
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(os.environ, REPO=helpers.ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    res = json.loads(line[len("RESULT "):])
    names = sorted(synth_es.CASES)[:6]
    expect, total = {}, 0
    for name in names:
        fr, idx, d = helpers.decode_all(helpers.oracle_lib(), [(0.0, synth_es.make_case(name))])
        total += len(fr)
        h = hashlib.sha256()
        for planes in fr:
            for p in planes:
                h.update(p.tobytes())
        expect[name] = h.hexdigest()
        d.destroy()
    assert res["digests"] == expect
    assert res["total"] == total
    assert res["seconds"] == 2.0  # max over ranks
