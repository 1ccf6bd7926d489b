"""CPU (gloo, world_size 2) test of the N>1 host-side logic: the distributed plan protocol
partial-aggregate -> hash exchange on the keys (Spark Murmur3 pmod) -> final-aggregate, and the
SinglePartition exchange of keyless aggregates, reproduce the single-process result; plus the
`bench.py --impl reference` contract under torchrun (rank 0 prints, the others exit 0)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from oracle import spark_cpu as O, spark_hash as H
from tests import datagen as G
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
def shard(r):
    rng = np.random.default_rng(77 + r)
    n = 3000 + 500 * r
    return [G.gen_column(rng, (O.INT64, 0, 0), n, null_frac=0.05, distinct=200), G.gen_column(rng, (O.DECIMAL64, 12, 2), n, null_frac=0.1)]
specs = [(O.AGG_SUM, 1, O.DECIMAL128, 2, 22), (O.AGG_COUNT, 1), (O.AGG_COUNT_ALL, 0)]
merge = [(O.AGG_SUM, 1, O.DECIMAL128, 2, 22), (O.AGG_SUM, 2, O.INT64, 0, 0), (O.AGG_SUM, 3, O.INT64, 0, 0)]
partial = O.groupby_cols(shard(rank), [0], specs)                      # GpuHashAggregateExec partial
cols, offs = H.hash_partition(partial, [0], world)                     # GpuHashPartitioning (seed 42, pmod world)
send = [[ (c.values[offs[r]:offs[r+1]], c.valid[offs[r]:offs[r+1]]) for c in cols] for r in range(world)]
recv = [None] * world
dist.all_to_all_object = None
out = [None] * world
for r in range(world):                                                 # all-to-all as world scatters (gloo has no all_to_all for objects)
    lst = [None]
    dist.scatter_object_list(lst, send if rank == r else None, src=r)
    out[r] = lst[0]
got = [O.OCol(np.concatenate([o[i][0] for o in out]), np.concatenate([o[i][1] for o in out]), cols[i].typ) for i in range(len(cols))]
final = O.groupby_cols(got, [0], merge)                                # GpuHashAggregateExec final (merge aggregates)
rows = O.rows_of(final)
# keys must be co-located: every key on exactly one rank
allrows = [None] * world
dist.all_gather_object(allrows, rows)
# keyless: SinglePartition exchange to rank 0
red = O.rows_of(O.reduce_cols(shard(rank), [(O.AGG_SUM, 1, O.DECIMAL128, 2, 22), (O.AGG_COUNT, 1)]))
allred = [None] * world
dist.gather_object(red, allred if rank == 0 else None, dst=0)
if rank == 0:
    merged = [r for rs in allrows for r in rs]
    keys = [r[0] for r in merged]
    assert len(keys) == len(set(keys)), "a group landed on two ranks"
    every = [shard(r) for r in range(world)]
    cat = [O.OCol(np.concatenate([s[i].values for s in every]), np.concatenate([s[i].valid for s in every]), every[0][i].typ) for i in range(2)]
    exp = O.rows_of(O.groupby_cols(cat, [0], specs))
    assert G.norm_rows(merged) == G.norm_rows(exp)
    tot = sum(r[0][0] or 0 for r in allred); cnt = sum(r[0][1] for r in allred)
    e = O.rows_of(O.reduce_cols(cat, [(O.AGG_SUM, 1, O.DECIMAL128, 2, 22), (O.AGG_COUNT, 1)]))[0]
    assert (tot, cnt) == e
    print("DIST_OK", len(merged))
dist.barrier()
dist.destroy_process_group()
'''


Q3_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
import bench                                    # the product bench's host-side sharding (q3_host_chunks / q3_chunks_of_rank)
from benchdata import tpch as gen
from oracle import spark_cpu as O, spark_hash as H, tpch as T
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sf = 0.002
chunks = bench.q3_host_chunks(sf, rank, world)   # this rank's 1/world of every table, as bench.py shards it
def cat(table, col):
    parts = [c[col] for c in chunks[table]]
    return np.concatenate(parts) if parts else np.zeros(0, np.int64)
def exchange(cols, key):                         # GpuShuffleExchangeExec: Spark Murmur3 pmod world on `key`, all-to-all
    kc = O.OCol(cols[key].astype(np.int64), np.ones(len(cols[key]), bool), (O.INT64, 0, 0))
    pid = H.partition_ids([kc], world) if len(cols[key]) else np.zeros(0, np.int32)
    send = [{k: v[pid == r] for k, v in cols.items()} for r in range(world)]
    got = []
    for r in range(world):
        lst = [None]
        dist.scatter_object_list(lst, send if rank == r else None, src=r)
        got.append(lst[0])
    return {k: np.concatenate([g[k] for g in got]) for k in cols}
# filters (each rank on its own slice)
ck = []
for c in chunks["customer"]:
    ck.append(c["c_custkey"][c["c_mktsegment_code"] == gen.SEGMENTS.index(gen.Q3_SEGMENT)])
cust = exchange({"c_custkey": np.concatenate(ck) if ck else np.zeros(0, np.int64)}, "c_custkey")
om = cat("orders", "o_orderdate") < gen.Q3_DATE
orders = exchange({k: cat("orders", k)[om] for k in ("o_orderkey", "o_custkey", "o_orderdate", "o_shippriority")}, "o_custkey")
hit = np.isin(orders["o_custkey"], cust["c_custkey"])           # co-partitioned: the join is local
j1 = exchange({k: orders[k][hit] for k in ("o_orderkey", "o_orderdate", "o_shippriority")}, "o_orderkey")
lm = cat("lineitem", "l_shipdate") > gen.Q3_DATE
line = exchange({k: cat("lineitem", k)[lm] for k in ("l_orderkey", "l_extendedprice", "l_discount")}, "l_orderkey")
info = {int(k): (int(d), int(p)) for k, d, p in zip(j1["o_orderkey"], j1["o_orderdate"], j1["o_shippriority"])}
rev = {}
for k, p, d in zip(line["l_orderkey"], line["l_extendedprice"], line["l_discount"]):
    if int(k) in info:
        rev[int(k)] = rev.get(int(k), 0) + int(p) * (100 - int(d))
local = sorted(((k, v) + info[k] for k, v in rev.items()), key=lambda r: (-r[1], r[2], r[0]))[:10]   # GpuTopN per rank
alltop = [None] * world
dist.gather_object(local, alltop if rank == 0 else None, dst=0)                                       # SinglePartition exchange
if rank == 0:
    final = sorted([r for t in alltop for r in t], key=lambda r: (-r[1], r[2], r[0]))[:10]
    assert final == T.q3_expected(sf, 42, threads=2), (final,)
    owned = [set(gen.q3_chunks_of_rank("lineitem", r, world)) for r in range(world)]
    assert set().union(*owned) == set(range(gen.Q3_CHUNKS["lineitem"])) and sum(len(o) for o in owned) == gen.Q3_CHUNKS["lineitem"]
    print("Q3_DIST_OK", final[0])
dist.barrier()
dist.destroy_process_group()
'''


# GpuShuffleExchangeExec's termination protocol (csrc/exec.cu: one batch of look-ahead, `more` flag in the exchange header,
# b2_comm_set_more / b2_comm_any_more) restated over gloo: every rank keeps calling while ANY rank has more, the call that
# carries a rank's last batch says so, and no extra empty round is spent.  Ranks with different batch counts (one with none)
# must agree on the number of rounds and lose no rows.
MORE_WORKER = r'''
import os, sys
import torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for case, counts in enumerate([(3, 1), (0, 2), (2, 2), (0, 0), (1, 4)]):
    batches = [[(rank, b, i) for i in range(5 + b)] for b in range(counts[rank])]       # this rank's child output
    it = iter(batches)
    ahead = next(it, None)                                                              # primed look-ahead
    rounds, received, finished = 0, [], False
    while not finished:
        cur, ahead = ahead, (next(it, None) if ahead is not None else None)
        hdr = {"has_data": cur is not None, "more": ahead is not None}                 # b2_comm_set_more(ahead != null)
        hdrs = [None] * world
        dist.all_gather_object(hdrs, hdr)                                               # the header all-gather of the call
        rows = [None] * world
        dist.all_gather_object(rows, cur or [])                                         # the data (every rank sees every row here)
        rounds += 1
        any_data, any_more = any(h["has_data"] for h in hdrs), any(h["more"] for h in hdrs)
        if any_data:
            received += [r for part in rows for r in part]
        finished = not any_more                                                         # b2_comm_any_more == 0: no further round
    expect_rounds = max(1, max(counts))                                                 # never an extra empty round
    assert rounds == expect_rounds, (case, rank, rounds, expect_rounds)
    want = sorted((r, b, i) for r in range(world) for b in range(counts[r]) for i in range(5 + b))
    assert sorted(received) == want, (case, rank)
print("MORE_OK")
dist.barrier()
dist.destroy_process_group()
'''


def _torchrun(args, script_args=(), timeout=300):
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29531"] + list(args) + list(script_args), capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_partial_exchange_final_protocol_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


def test_exchange_lookahead_termination_protocol_world2(tmp_path):
    script = tmp_path / "more_worker.py"
    script.write_text(MORE_WORKER)
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("MORE_OK") == 2


def test_q3_strong_scaled_plan_protocol_world2(tmp_path):
    """the bench's strong-scaled q3 (its own host-side sharding: bench.q3_host_chunks) with the exchanges emulated over gloo:
    filters on each rank's slice, Murmur3 exchanges on the join keys, local joins, local top-10, final top-10 == numpy q3"""
    script = tmp_path / "q3_worker.py"
    script.write_text(Q3_WORKER % {"root": ROOT})
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "Q3_DIST_OK" in r.stdout


def test_reference_arm_under_torchrun_prints_once():
    env_rows = ["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--ref-sf", "0.1"]
    r = _torchrun([os.path.join(ROOT, "bench.py")], env_rows)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rows/s" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
