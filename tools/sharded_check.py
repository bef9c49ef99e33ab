"""Parity of the segment-sharded search behind the C ABI (dgpu_sharded_*): run under torchrun with one rank per GPU.
Every rank opens its run of segments of a scaled C2 corpus, the ranks join an NCCL communicator through the library
(torch.distributed only ships the 128-byte id), and the merged results of dgpu_sharded_search_batch_text must equal,
bit for bit, those of one reader holding the whole corpus (opened on rank 0's GPU by every rank in turn would cost
memory: rank 0 computes the expectation and broadcasts it). Exit code 0 = identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diagon_b200 as dg  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")   # plumbing for the id and the expectation only
    scale = float(os.environ.get("DGPU_CHECK_SCALE", "0.02"))
    failures = 0
    for shape, kind, k, n in (("C2", "OR body 0", 10, 3000), ("C3-AND2", "AND body", 10, 500), ("C4", "ORF body price", 100, 500),
                              ("C5", "OR body 0", 1000, 64)):
        spec = dg.named_corpus("C4", scale)
        segs = spec.num_segments
        lo, hi = segs * rank // world, segs * (rank + 1) // world
        reader = dg.IndexReader.synthetic(spec, local, lo, hi)
        uid = [dg.ShardedSearcher.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ss = dg.ShardedSearcher(reader, uid[0], rank, world)
        text = dg.query_log_text(shape, spec.vocab, n, kind)
        for chunks in (1, 3, 0):   # 0: a stream of submitted batches, three in flight, the middle one is checked
            if chunks:
                reader.set_option("pipeline_chunks", chunks)
                reader.set_option("pipeline_min", 1 if chunks > 1 else 1 << 30)
                got = ss.search_batch_text(text, k)
            else:
                half = text[: text.index(b"\n", len(text) // 2) + 1]
                tickets = [ss.submit_batch_text(half, k), ss.submit_batch_text(text, k), ss.submit_batch_text(half, k)]
                tickets[0].collect()
                got = tickets[1].collect()
                tickets[2].collect()
            want = [None]
            if rank == 0:
                whole = dg.IndexReader.synthetic(spec, local)
                w = dg.IndexSearcher(whole).search_batch_text(text, k)
                want = [(w.docs.copy(), w.scores.copy(), w.counts.copy(), w.total_hits.copy())]
                whole.close()
            dist.broadcast_object_list(want, src=0)
            wd, wsc, wc, wh = want[0]
            ok = np.array_equal(got.total_hits, wh) and np.array_equal(got.counts, wc)
            for q in range(len(wc)):
                c = int(wc[q])
                ok = ok and np.array_equal(got.docs[q, :c], wd[q, :c]) and np.array_equal(got.scores[q, :c], wsc[q, :c])
            if not ok:
                failures += 1
            if rank == 0:
                print(f"sharded {shape} k={k} chunks={chunks} world={world}: {'identical' if ok else 'MISMATCH'} "
                      f"({int(wh.sum())} hits over {len(wc)} queries)", flush=True)
        ss.close()
        reader.close()
    t = torch.tensor([failures])
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(1 if int(t[0]) else 0)


if __name__ == "__main__":
    main()
