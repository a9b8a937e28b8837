"""Run the ranks of a multi-GPU job as THREADS of this process on ONE device (library backend "LOCAL:<name>", csrc/comm.cu):
the partition, ghost layers, halo plans, staged peer-memory exchanges and the distributed multigrid hierarchy run exactly as
with one process per GPU, only the transport differs -- so `pytest -m gpu` on a single-GPU box covers the N>1 logic."""
import itertools
import threading

_counter = itertools.count()


def run_ranks(n, fn, device=0):
    """fn(rank, comm) in n threads; comm = (id, rank, n) for ddps_b200.Session(device, comm=comm).  Returns the list of
    results in rank order; the first exception of any rank is re-raised (the others then fail on the broken group)."""
    ident = (b"LOCAL:test-%d" % next(_counter)).ljust(128, b"\0")
    out, err = [None] * n, [None] * n

    def work(r):
        try:
            out[r] = fn(r, (ident, r, n))
        except BaseException as e:   # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None and "stopped participating" not in str(e):
            raise e
    for e in err:
        if e is not None:
            raise e
    return out
