"""The N > 1 path on CPU: two `gloo` ranks, segments sharded 4 + 4.

What runs on the host in a sharded search is (1) the statistics exchange that makes idf / avgdl identical on every
rank (TermQuery.cpp:184-260 uses global numbers), and (2) the protocol around the single exchange step: local top-k
lists as 64-bit keys, one all_gather, a k-way merge, totalHits = sum of the local counts. Both are exercised here
with host-only readers (no engine: nothing is scored by the product on CPU — the local lists come from the oracle,
which is the checker). The CUDA merge kernel itself is covered by tests/test_gpu_parity.py.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K = 10
N_QUERIES = 40


def _keys(docs_scores):
    """TopDocs -> descending 64-bit keys, the engine's result encoding (include/dgpu_engine.h: dgpu_results)."""
    out = np.zeros(K, dtype=np.uint64)
    for i, (d, s) in enumerate(docs_scores[:K]):
        b = int(np.float32(s).view(np.uint32))
        o = (~b & 0xFFFFFFFF) if b & 0x80000000 else (b | 0x80000000)
        out[i] = (o << 32) | (0xFFFFFFFF - d)
    return out


def _merge(all_keys, all_counts):
    """numpy mirror of merge_parts_kernel: best K of the union, keys are unique."""
    cat = np.concatenate([all_keys[p][: all_counts[p]] for p in range(len(all_keys))])
    return np.sort(cat)[::-1][:K]


def _worker(rank, world, port, tmp):
    import torch.distributed as dist

    import diagon_b200 as dg
    from diagon_b200 import api
    from oracle import oracle as orc

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        import torch

        spec = dg.named_corpus("C2", 0.002)
        nseg = spec.num_segments
        lo, hi = nseg * rank // world, nseg * (rank + 1) // world
        local = dg.IndexReader.synthetic(spec, -1, lo, hi)   # host-only: dictionary + statistics, no engine
        whole = dg.IndexReader.synthetic(spec, -1)

        # (1) statistics exchange
        df = torch.from_numpy(local.get_doc_freqs().copy())
        dist.all_reduce(df)
        ttf, md = local.get_field_totals("body")
        tot = torch.tensor([ttf, md], dtype=torch.int64)
        dist.all_reduce(tot)
        assert np.array_equal(df.numpy(), whole.get_doc_freqs()), "global docFreq differs from the unsharded index"
        assert (int(tot[0]), int(tot[1])) == whole.get_field_totals("body")
        local.set_doc_freqs(df.numpy())
        local.set_field_totals("body", int(tot[0]), int(tot[1]))
        assert np.array_equal(local.get_doc_freqs(), whole.get_doc_freqs())

        # (2) gather + merge of local top-k lists. Local list of a shard = the best K among the hits whose doc lies in
        # the shard's segments, scored with GLOBAL statistics: exactly what a rank's engine returns.
        path = os.path.join(tmp, "c2s.dmp")
        if rank == 0:
            dg.write_synthetic_dump(spec, path)
        dist.barrier()
        dump = dg.read_dump(path)
        ox = orc.OracleIndex(dump)
        doc_lo = dump.segments[lo].doc_base
        doc_hi = dump.segments[hi - 1].doc_base + dump.segments[hi - 1].max_doc
        lines = dg.query_log_text("C2", spec.vocab, N_QUERIES, "OR body 0").decode().strip().split("\n")
        for line in lines:
            q = api.parse_line(line)
            hits_all, everything, _ = ox.search(q, dump.max_doc)          # every hit, best first
            mine = [(d, s) for d, s in everything if doc_lo <= d < doc_hi]
            keys = torch.from_numpy(_keys(mine).view(np.int64).copy())
            meta = torch.tensor([min(len(mine), K), len(mine)], dtype=torch.int64)
            g_keys = [torch.zeros(K, dtype=torch.int64) for _ in range(world)]
            g_meta = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(g_keys, keys)
            dist.all_gather(g_meta, meta)
            merged = _merge([k.numpy().view(np.uint64) for k in g_keys], [int(m[0]) for m in g_meta])
            want_hits, want, _ = ox.search(q, K)
            assert sum(int(m[1]) for m in g_meta) == want_hits == hits_all, line
            assert np.array_equal(merged, _keys(want)[: len(want)]), line
        # (3) the host work of a batch divided between the ranks: every rank compiles its slice of the query lines
        # into a relocatable blob, one all_gather exchanges them, and the concatenation is the batch a single rank
        # would have compiled (descriptors depend on global statistics only)
        ls = dg.IndexSearcher(local)
        ws = dg.IndexSearcher(whole)
        n = len(lines)
        mine_text = ("\n".join(lines[n * rank // world: n * (rank + 1) // world]) + "\n").encode()
        blob = ls.compile_batch_text(mine_text)
        size = torch.tensor([blob.size], dtype=torch.int64)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, size)
        cap = max(int(x[0]) for x in sizes)
        padded = torch.zeros(cap, dtype=torch.uint8)
        padded[: blob.size] = torch.from_numpy(blob)
        gathered = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(gathered, padded)
        blobs = [g.numpy()[: int(sz[0])] for g, sz in zip(gathered, sizes)]

        def unpack(b):
            magic, nq, nt, nf = np.frombuffer(b[:16].tobytes(), dtype=np.uint32)
            assert magic == 0x42504744
            q = np.frombuffer(b[16:16 + 24 * nq].tobytes(), dtype=np.uint32).reshape(nq, 6).copy()   # sizeof(dgpu_query) == 24
            t = b[16 + 24 * nq: 16 + 24 * nq + 12 * nt].tobytes()
            return q, t, nt, nf

        want_q, want_t, _, _ = unpack(ws.compile_batch_text(("\n".join(lines) + "\n").encode()))
        got_q, got_t, toff = [], b"", 0
        for b in blobs:
            q, t, nt, nf = unpack(b)
            assert nf == 0
            q[:, 0] += toff
            q[:, 1] += toff
            got_q.append(q)
            got_t += t
            toff += nt
        assert np.array_equal(np.concatenate(got_q), want_q) and got_t == want_t
        ls.close()
        ws.close()
        local.close()
        whole.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_statistics_exchange_and_topk_merge(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def _shm_worker(rank, world, uid, q):
    from diagon_b200 import _lib
    import ctypes as C

    buf = (C.c_uint8 * 128).from_buffer_copy(uid)
    rc = _lib.load().dgpu_shm_exchange_selftest(C.addressof(buf), rank, world, 50)
    q.put((rank, rc, _lib.last_error() if rc else ""))


@pytest.mark.parametrize("world", [2, 3])
def test_shared_memory_compile_exchange(world):
    """The channel through which the ranks of one box swap the compiled slices of a batch (diagon_b200/host/shm_exchange.h):
    `world` processes, 50 rounds of patterned slices of varying size, double-buffered; every slice must arrive intact."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    uid = bytes((i * 37 + 11 * world) & 0xFF for i in range(128))
    q = ctx.Queue()
    procs = [ctx.Process(target=_shm_worker, args=(r, world, uid, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r for r, _, _ in got) == list(range(world))
    assert all(rc == 0 for _, rc, _ in got), got
