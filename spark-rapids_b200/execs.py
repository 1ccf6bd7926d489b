"""Python face of the C++ operator layer (csrc/exec.cu): same node names as the reference's GpuExec
classes, same pull-iterator-of-batches contract (GpuExec.internalDoExecuteColumnar)."""
import ctypes

from ._lib import lib, check
from . import _init as m


class GpuExec:
    def __init__(self, handle, keep=()):
        self.h = ctypes.c_int64(handle)
        self._keep = list(keep)  # programs / children / host buffers that must outlive the node

    def __del__(self):
        if getattr(self, "h", None) is not None and self.h.value:
            lib.b2_exec_close(self.h)
            self.h = ctypes.c_int64(0)

    def next(self):
        out = ctypes.c_int64()
        check(lib.b2_exec_next(self.h, ctypes.byref(out)))
        return m.Table(out.value) if out.value else None

    def __iter__(self):
        while True:
            t = self.next()
            if t is None:
                return
            yield t

    def collect(self):
        """all output batches concatenated (like executeCollect on the columnar side)"""
        batches = list(self)
        if not batches:
            return None
        return batches[0] if len(batches) == 1 else m.concat(batches)

    @property
    def metrics(self):
        out = (ctypes.c_int64 * 3)()
        check(lib.b2_exec_metrics(self.h, out))
        return {"numOutputRows": out[0], "numOutputBatches": out[1], "opTime": out[2]}


def _new(fn, *args, keep=()):
    out = ctypes.c_int64()
    check(fn(*args, ctypes.byref(out)))
    return GpuExec(out.value, keep)


def GpuBatchSource(tables=()):
    e = _new(lib.b2_exec_source)
    for t in tables:
        check(lib.b2_exec_source_push(e.h, t.h))
    return e


def GpuParquetScanExec(buffers, columns):
    import numpy as np
    names = (ctypes.c_char_p * len(columns))(*[c.encode() for c in columns])
    e = _new(lib.b2_exec_parquet_scan, names, len(columns))
    for b in buffers:
        arr = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b
        e._keep.append(arr)
        check(lib.b2_exec_parquet_scan_add(e.h, m._ptr(arr), arr.nbytes))
    return e


def GpuFilterExec(condition, child):
    prog = m.Program([condition])
    return _new(lib.b2_exec_filter, child.h, prog.h, keep=[prog, child])


def GpuProjectExec(project_list, child):
    prog = m.Program(project_list)
    return _new(lib.b2_exec_project, child.h, prog.h, keep=[prog, child])


def GpuHashAggregateExec(child, grouping, aggregates, pre_project=None, condition=None, mode="partial"):
    """mode 'partial'/'complete': update aggregates over `pre_project` expressions (with `condition`
    fused in as the child filter); 'final': merge aggregation buffers whose keys lead the input."""
    if mode == "final":
        return _new(lib.b2_exec_hash_aggregate, child.h, ctypes.c_int64(0), 0, 1, m._i32s(grouping), len(grouping), m._agg_specs(aggregates),
                    len(aggregates), keep=[child])
    exprs = ([condition] if condition is not None else []) + list(pre_project)
    prog = m.Program(exprs)
    return _new(lib.b2_exec_hash_aggregate, child.h, prog.h, int(condition is not None), 0, m._i32s(grouping), len(grouping),
                m._agg_specs(aggregates), len(aggregates), keep=[prog, child])


def GpuShuffledHashJoinExec(stream_keys, build_keys, join_type, stream, build, nulls_equal=False):
    return _new(lib.b2_exec_shuffled_hash_join, stream.h, build.h, m._i32s(stream_keys), m._i32s(build_keys), len(stream_keys), join_type,
                int(nulls_equal), keep=[stream, build])


def GpuSortExec(sort_order, child, global_sort=True):
    return _new(lib.b2_exec_sort, child.h, m._order_args(sort_order), len(sort_order), int(global_sort), -1, keep=[child])


def GpuTopN(limit, sort_order, child):
    return _new(lib.b2_exec_sort, child.h, m._order_args(sort_order), len(sort_order), 1, limit, keep=[child])


def GpuCoalesceBatches(child, target_rows):
    return _new(lib.b2_exec_coalesce, child.h, target_rows, keep=[child])


def GpuShuffleExchangeExec(child, key_cols, comm=None, world=1):
    return _new(lib.b2_exec_shuffle_exchange, child.h, m._i32s(key_cols), len(key_cols), comm.h if comm else ctypes.c_int64(0), world,
                keep=[child, comm])
