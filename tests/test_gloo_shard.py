"""world_size-2 gloo tests of the multi-GPU host logic: the per-wave exchange (rlen chain, output offsets) and the whole
launcher (`python -m minialign_b200.mgpu`) run on two ranks through the emulation library, merged SAM compared with the
one-process run."""
import os
import socket
import subprocess

import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLD, build_emu, build_emu_cli
from minialign_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _exchange_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = shard.WaveExchange(host_group=dist.new_group(backend="gloo"))
    log = []
    # wave 0: rank 0 leaves 700 behind, rank 1 loads no chain; wave 1: rank 0's value moves when it is committed (700 -> 650 -> 640),
    # rank 1 must see the corrected one; wave 2: partial (rank 1 has no chunk)
    script = {0: [(True, 700), (True, 650), (True, 300)], 1: [(False, 0), (True, 900), (False, 0)]}
    for w in range(3):
        valid, value = script[rank][w]
        ex.begin_wave(valid, value)
        seen = []

        def commit(v, w=w, valid=valid, value=value):
            seen.append(v)
            if w == 1 and rank == 0:
                return True, 640                     # the committed chunk leaves something else behind than it announced
            return valid, value
        ex.settle(commit)
        ofs, total = ex.offsets(100 * (w + 1) + rank if (w < 2 or rank == 0) else 0)
        log.append((seen, ex.rlen, ofs, total))
    q.put((rank, log))
    dist.barrier()
    dist.destroy_process_group()


def test_wave_exchange_chain_and_offsets():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    out = dict(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    r0, r1 = out[0], out[1]
    # wave 0: both start from the fresh thread's 0; rank 1 (after rank 0 in chunk order) sees 700; the wave leaves 700
    assert r0[0][0] == [0] and r1[0][0] == [700] and r0[0][1] == r1[0][1] == 700
    # wave 1: rank 0 committed with 700, announced 650 but leaves 640: a second round hands rank 1 the corrected value
    assert r0[1][0] == [700, 700] and r1[1][0] == [650, 640] and r0[1][1] == r1[1][1] == 900
    # wave 2: partial wave
    assert r0[2][0] == [900] and r0[2][1] == r1[2][1] == 300
    # offsets: exclusive prefix in rank order, carried from wave to wave
    assert (r0[0][2], r1[0][2]) == (0, 100) and r0[0][3] == r1[0][3] == 201
    assert (r0[1][2], r1[1][2]) == (201, 401) and r0[1][3] == 401
    assert r0[2][2] == 602 and r0[2][3] == r1[2][3] == 300


def test_single_process_is_identity():
    ex = shard.WaveExchange()
    ex.begin_wave(True, 5)
    assert ex.settle(lambda v: (True, 5)) == 0 and ex.rlen == 5
    assert ex.offsets(123) == (0, 123) and ex.offsets(7) == (123, 7)
    assert shard.deal_chunks(5, 1, 0) == [0, 1, 2, 3, 4] and shard.deal_chunks(7, 2, 1) == [1, 3, 5]


def test_mgpu_two_ranks_merged_sam_equals_one_process(tmp_path):
    """The launcher on two gloo ranks (emulation library): chunks dealt round-robin, rlen chained across ranks, text written
    with pwrite at exchanged offsets -> one SAM, identical to the single-process CLI's."""
    so, cli = build_emu(), build_emu_cli()
    lines = open(os.path.join(GOLD, "reads.fa")).read().split("\n")
    recs = [(lines[i], lines[i + 1]) for i in range(0, len(lines) - 1, 2) if len(lines[i + 1]) <= 2500][:36]
    fa = str(tmp_path / "sub.fa")
    with open(fa, "w") as f:
        for h, s in recs:
            f.write(h + "\n" + s + "\n")
    idx = os.path.join(GOLD, "small.mai")
    one = subprocess.run([cli, "-xpacbio", "-TAS,NM,MD,SA", "-c1", idx, fa], capture_output=True)
    assert one.returncode == 0
    exp = [l for l in one.stdout.decode().split("\n") if not l.startswith("@PG")]
    out = str(tmp_path / "merged.sam")
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    p = subprocess.run(["python", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        "-m", "minialign_b200.mgpu", "--backend", "gloo", "--lib", so, "-xpacbio", "-TAS,NM,MD,SA", "-c2", "-N0.01", "-o", out, idx, fa],
                       capture_output=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-1500:]
    got = [l for l in open(out).read().split("\n") if not l.startswith("@PG")]
    assert got == exp
