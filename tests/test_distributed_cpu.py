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


def _torchrun(args, script_args=(), timeout=300):
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29531"] + list(args) + list(script_args), capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_partial_exchange_final_protocol_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout


def test_reference_arm_under_torchrun_prints_once():
    env_rows = ["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--rows", "200000"]
    r = _torchrun([os.path.join(ROOT, "bench.py")], env_rows)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rows/s" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
