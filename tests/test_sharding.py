"""Multi-GPU host logic on CPU (gloo, world_size 2): reads are sharded by anchors across ranks, every rank chains only its
shard (here: the oracle stands in for the device, this is the host-side partition + reduction that is under test) and the
whole-job figures are the SUM of the work and the MAX of the time over ranks -- no collective on the data path."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import __graft_entry__ as entry
pkg = entry.load_package()
po = entry.load_oracle()
from mm2gb_b200 import synth, sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
a, off = synth.ont_like_batch(seed=9, n_reads=37, lo=20, hi=900)
a_r, off_r, r0, r1 = sharding.shard(a, off, world, rank)
prm = po.map_ont_params()
pairs, _, _ = po.lchain_batch(prm, a_r, off_r, 0, len(off_r) - 1, n_threads=1)
sums, maxes = sharding.reduce_job(dist, [pairs, int(off_r[-1]), r1 - r0], [10.0 + rank])
gathered = [None] * world
dist.all_gather_object(gathered, (r0, r1))
if rank == 0:
    whole, _, _ = po.lchain_batch(prm, a, off, 0, len(off) - 1, n_threads=1)
    json.dump({"sums": sums, "maxes": maxes, "whole_pairs": int(whole), "anchors": int(off[-1]), "reads": len(off) - 1,
               "ranges": gathered}, open(sys.argv[2], "w"))
dist.destroy_process_group()
'''


def test_shard_bounds_properties(pkg):
    from mm2gb_b200 import sharding
    rng = np.random.default_rng(3)
    for n_reads in (0, 1, 2, 7, 100, 1000):
        n = rng.integers(0, 5000, n_reads)
        off = np.zeros(n_reads + 1, np.int64)
        off[1:] = np.cumsum(n)
        for world in (1, 2, 4, 8):
            b = sharding.shard_bounds(off, world)
            assert b[0] == 0 and b[-1] == n_reads and len(b) == world + 1 and np.all(np.diff(b) >= 0)
            if n_reads >= 100:   # balanced by anchors to within one (largest) read
                per = np.diff(off[b])
                assert per.max() - per.min() <= 2 * n.max(), (n_reads, world, per)
    a = np.arange(20, dtype=np.uint64).reshape(10, 2)
    off = np.array([0, 3, 3, 7, 10], np.int64)
    parts = [sharding.shard(a, off, 2, r) for r in range(2)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), a)
    assert all(p[1][0] == 0 for p in parts) and parts[0][3] == parts[1][2]


def test_two_ranks_gloo_sum_of_work_max_of_time(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "out.json"
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(out)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        txt, _ = p.communicate(timeout=300)
        assert p.returncode == 0, txt.decode()[-2000:]
    res = json.load(open(out))
    assert res["sums"][0] == res["whole_pairs"] and res["sums"][1] == res["anchors"] and res["sums"][2] == res["reads"]
    assert res["maxes"] == [11.0]
    (a0, a1), (b0, b1) = res["ranges"]
    assert a0 == 0 and a1 == b0 and b1 == res["reads"]


def test_rank_placement_plan(pkg):
    """host placement of a rank (bench under torchrun): CPUs of the GPU's NUMA node, or nothing where no topology shows"""
    from mm2gb_b200 import sharding
    allc = list(range(64))
    node0, node1 = list(range(0, 32)), list(range(32, 64))
    gpus = [node0] * 4 + [node1] * 4
    for r in range(8):
        cpus, sharers = sharding.plan_rank_cpus(gpus, r, allc)
        assert cpus == (node0 if r < 4 else node1) and sharers == 4
    # affinity mask of the process narrower than the node
    cpus, sharers = sharding.plan_rank_cpus(gpus, 5, list(range(16, 48)))
    assert cpus == list(range(32, 48)) and sharers == 4
    # flat box: every GPU reports every CPU -> leave the scheduler alone
    assert sharding.plan_rank_cpus([allc] * 8, 3, allc) == (None, 8)
    # NVML silent for one GPU, or a node without usable CPUs -> no pinning either
    assert sharding.plan_rank_cpus([node0, []], 0, allc) == (None, 2)
    assert sharding.plan_rank_cpus([node0, node1], 1, node0) == (None, 2)
    # uneven: 3 GPUs on one node, 1 on the other
    cpus, sharers = sharding.plan_rank_cpus([node0, node0, node0, node1], 3, allc)
    assert cpus == node1 and sharers == 1
    # one rank / no NVML in this container: a report, never an exception, and the affinity is untouched
    import os
    before = os.sched_getaffinity(0)
    rep = sharding.place_rank(0, 1)
    assert rep["pinned"] is False
    rep = sharding.place_rank(1, 2)
    assert rep["pinned"] is False and os.sched_getaffinity(0) == before


def test_cpulist_parser_of_the_staging_placement(tmp_path):
    """csrc/host_place.h: the sysfs local_cpulist format ("0-31,64-95") -> cpu set (host-only code, compiled with g++)"""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include "host_place.h"
#include <cstring>
int main(int argc, char **argv) {
    cpu_set_t s;
    int n = mm2gb::parse_cpulist(argv[1], &s);
    printf("%d", n);
    for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &s)) printf(" %d", c);
    printf("\n");
    return 0;
}
''')
    exe = tmp_path / "t"
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "mm2-gb_b200", "csrc"), "-I", cuda_inc, "-o", str(exe), str(src)],
                   check=True)

    def parse(s):
        out = subprocess.run([str(exe), s], check=True, capture_output=True, text=True).stdout.split()
        return int(out[0]), [int(x) for x in out[1:]]
    assert parse("0-3,8-9\n") == (6, [0, 1, 2, 3, 8, 9])
    assert parse("5") == (1, [5])
    assert parse("0-1,1-2") == (3, [0, 1, 2])
    assert parse("") == (0, [])
    assert parse("\n") == (0, [])
    assert parse("7-3") == (0, [])          # malformed range: nothing
    assert parse("2,x") == (1, [2])
