#!/usr/bin/env python
"""bench.py — BM25 top-k queries/sec of the B200 engine on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2] [--scale S]

One "step" = one pass of the hot path over one batch of synthetic queries (C2: 10,000 OR-10 queries,
top-10, MS-MARCO-passage-shaped corpus of 8,841,823 docs in 8 segments). Prints ONE JSON line:

  value     whole-job queries/s with the batch already staged on the device (kernels only; for N > 1 also the
            NCCL all-gather of the local top-k and the device merge), CUDA events, max over ranks
  e2e       the same through the reference-facing C ABI with HOST buffers, every step: query text in host memory ->
            parse, dictionary lookups, weights, H2D, kernels, D2H -> host arrays. N == 1: a stream of batches through
            dgpu_submit_batch_text / dgpu_collect_batch with two batches in flight (the reference arm is a throughput
            run too: 16 threads pulling queries); `one_call_at_a_time` inside it is dgpu_search_batch_text called
            K times in a row. N > 1: dgpu_sharded_search_batch_text on every rank
  roofline  algorithmic posting bytes of the batch / device time of the search kernel, against the measured
            HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, its own IndexSearcher, stock config) timed on this box's host cores on
            a bounded sample of the same workload (rank 0, N == 1 only)

--impl reference times only the reference arm and prints the same line shape with "impl": "reference".
Nothing here reads /root/reference; the reference binaries come prebuilt in oracle/_ref.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line. Libraries write there too (NCCL prints its version banner with printf), so
# file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a duplicate of the real stdout.
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

WORKLOADS = {
    # name: (corpus, query log, line prefix, k, default batch)
    "C2": ("C2", "C2", "OR body 0", 10, 10000),
    "C3-AND2": ("C2", "C3-AND2", "AND body", 10, 10000),
    "C3-AND4": ("C2", "C3-AND4", "AND body", 10, 10000),
    "C4": ("C4", "C4", "ORF body price", 100, 10000),
    "C5": ("C5", "C5", "OR body 0", 1000, 1000),
    "C1": ("C1", "C1", "OR body 0", 10, 1000),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="corpus scale (1.0 = the named configuration)")
    ap.add_argument("--queries", type=int, default=0, help="queries per batch (0 = the named batch size)")
    ap.add_argument("--log2-window", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--warps-per-sm", type=int, default=0)
    ap.add_argument("--window-docs", type=int, default=0)
    ap.add_argument("--stage-log2", type=int, default=0)
    ap.add_argument("--decode-ctas-per-sm", type=int, default=0)
    ap.add_argument("--part-factor", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=0, help="3 = batched decode_score + accumulate_topk (default), 2 = fused windows")
    ap.add_argument("--lane-merge", type=int, default=-1, help="1 = queries of <= 32 terms on staged_merge_topk_kernel (default), 2 = <= 16 terms on lane_merge_topk_kernel, 0 = windows")
    ap.add_argument("--lane-ctas-per-sm", type=int, default=0)
    ap.add_argument("--pool-smem-cap", type=int, default=0)
    ap.add_argument("--pipeline-chunks", type=int, default=0, help="chunks of the end-to-end text call (1 = no host/device overlap)")
    ap.add_argument("--lane-ring-entries", type=int, default=0)
    ap.add_argument("--union-window-docs", type=int, default=0, help="docs per window (bits of shared memory) of union_topk_kernel")
    ap.add_argument("--filter-stream", type=int, default=-1, help="0: range-filter values gathered per posting; 1 (default): streamed with the runs")
    ap.add_argument("--union-max-overlap", type=int, default=-1, help="percent of expected later sightings above which a query goes to staged_merge_topk_kernel")
    ap.add_argument("--cpu-sample-docs", type=int, default=200000)
    ap.add_argument("--cpu-sample-queries", type=int, default=0,
                    help="queries of the reference arm (0: --impl reference runs the whole batch, the cpu_baseline leg 400)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ reference arm
def reference_arm(args, corpus_name, log_name, kind, k, steps, warmup, full, batch):
    """Times the reference's own IndexSearcher (oracle/_ref, compiled from the unmodified sources) on this box's host
    cores: an index built by its own IndexWriter, all host threads (one DirectoryReader + IndexSearcher per thread; the
    reference's searcher is single-threaded and not thread-safe), the threads pulling queries from one counter.

    full=True (the --impl reference arm): the index covers as much of the corpus as the reference can index within
    DGPU_REF_INDEX_BUDGET_S seconds (default 200; the whole C2 corpus takes ~2 min on the GPU box's host) and is cached
    under /tmp for the following invocations on the same box; the WHOLE query batch is run (--cpu-sample-queries 0).
    full=False (the cpu_baseline leg of our own arm): the first `cpu_sample_docs` documents (or the cached index) and the
    first `cpu_sample_queries` queries. Nothing of the product is imported or loaded here."""
    from oracle import refrun

    if not refrun.driver():
        return None
    spec = refrun.corpus_spec(corpus_name, args.scale)
    n_docs, n_seg = spec["num_docs"], spec["num_segments"]
    nq = batch if (full and not args.cpu_sample_queries) else min(batch, args.cpu_sample_queries or 400)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    tmp = tempfile.mkdtemp(prefix="dgpu_ref_")
    price = corpus_name == "C4"
    try:
        best, best_idx = refrun.cached_index(corpus_name, args.scale)
        if full:
            budget = float(os.environ.get("DGPU_REF_INDEX_BUDGET_S", "200"))
            probe_docs = min(n_docs, 50000)
            rate = probe_docs / max(refrun.build_index(corpus_name, args.scale, probe_docs, 1, os.path.join(tmp, "probe"), price), 1e-3)
            docs = min(n_docs, max(probe_docs, int(rate * budget) // 1000 * 1000))
            if docs >= 0.85 * n_docs:   # close enough: index the whole corpus, so the two arms run the same config
                docs = n_docs
            if best >= 0.7 * docs:      # an index built by an earlier invocation on this box, about as large
                docs = best
        else:
            docs = max(min(args.cpu_sample_docs, n_docs), best)
        segments = n_seg if docs == n_docs else 1
        idx, index_s, reused = refrun.ensure_index(corpus_name, args.scale, docs, segments, price)
        index_note = (f"index reused from the box's cache (built in {index_s:.0f}s)" if reused
                      else f"indexed in {index_s:.0f}s by its own IndexWriter")
        qfile = os.path.join(tmp, "q.txt")
        refrun.write_queries(log_name, spec["vocab"], nq, kind, qfile)
        out = {}
        # one untimed pass (page cache, allocator) that also tells how many timed passes fit ~90 s; the exhaustive mode is
        # reported beside the headline: one pass is enough for it
        probe = refrun.search(idx, qfile, k, True, threads)
        passes = max(1, min(max(1, steps), int(90.0 / max(probe["search_seconds"], 1e-3))))
        out["default"] = refrun.search(idx, qfile, k, True, threads, warmup=0, repeat=passes)
        out["exhaustive"] = refrun.search(idx, qfile, k, False, threads)
        steps = passes
        frac = docs / n_docs
        return {
            # the headline is the reference's STOCK path: IndexSearcher::search(q, k) with its default config, i.e.
            # MaxScore / Block-Max WAND pruning on (hit counts are lower bounds there)
            "value": out["default"]["qps"], "unit": "queries/s", "cores": threads, "kind": "reference",
            "mode": "stock (enable_block_max_wand=true)",
            "sample": (f"reference IndexSearcher ({os.path.basename(refrun.driver())}), first {docs} of {n_docs} docs "
                       f"({100 * frac:.2f}% of the corpus, {segments} segment(s), {index_note}), "
                       f"{'all' if nq == batch else 'first'} {nq} of the batch's {batch} queries x {max(1, steps)} passes, "
                       f"{threads} threads (one reader + searcher each) pulling queries from one counter"),
            "queries": nq,
            # the same work as the GPU engine (exhaustive scoring, exact hit counts): enable_block_max_wand=false
            "exhaustive_mode_qps": out["exhaustive"]["qps"],
            "corpus_fraction": frac,
            "exhaustive_qps_scaled_to_full_corpus": out["exhaustive"]["qps"] * frac,
        }
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def parity_vs_reference(args, searcher, corpus_name, log_name, kind, k, spec, n_check=400):
    """Full-size parity: when this box holds the reference's index of the WHOLE corpus (the --impl reference arm ran
    first), the first `n_check` queries of the workload are run through the reference in exhaustive mode
    (enable_block_max_wand=false, IEEE build) and must match the GPU's results bit for bit. The oracle is the checker
    here, never the thing measured."""
    from oracle import refrun

    if not refrun.driver(False):
        return None
    docs, idx = refrun.cached_index(corpus_name, args.scale)
    if docs != spec.num_docs:
        return {"checked": 0, "note": "no whole-corpus reference index on this box (run --impl reference first)"}
    tmp = tempfile.mkdtemp(prefix="dgpu_par_")
    try:
        qfile, rfile = os.path.join(tmp, "q.txt"), os.path.join(tmp, "r.res")
        refrun.write_queries(log_name, spec.vocab, n_check, kind, qfile)
        refrun.search(idx, qfile, k, False, max(1, min(os.cpu_count() or 1, 64)), out=rfile, fast=False)
        _, hits, counts, docs_, scores = refrun.read_results(rfile)
        got = searcher.search_batch_text(open(qfile, "rb").read(), k)
        bad = refrun.compare(got, hits, counts, docs_, scores)
        return {"checked": int(len(hits)), "mismatches": len(bad), "first_mismatch": (bad[0] if bad else None),
                "against": "reference IndexSearcher, enable_block_max_wand=false, IEEE build, whole corpus",
                "total_hits_checked": int(hits.sum())}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------ main
def lane_kernel_name(bstats):
    return "lane_merge_topk_kernel" if int(bstats[11]) == 2 else "staged_merge_topk_kernel"


def dominant_kernel(bstats):
    """The scoring kernel most work items of the batch went to (they run one after the other in phase [1])."""
    items = {"accumulate_topk_kernel": int(bstats[6]), "intersect_topk_kernel": int(bstats[7]),
             lane_kernel_name(bstats): int(bstats[10]), "union_topk_kernel": int(bstats[13])}
    return max(items, key=items.get)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    corpus_name, log_name, kind, k, batch = WORKLOADS[args.workload]
    if args.queries:
        batch = args.queries

    def workload_text(spec):
        return (f"{args.workload}: {batch} x {kind.split()[0]}-{log_name} top-{k} on the {corpus_name} synthetic corpus "
                f"({spec.num_docs} docs, vocab {spec.vocab}, {spec.num_segments} segments, scale {args.scale})")

    if args.impl == "reference":
        if rank != 0:
            return 0
        # nothing of the product is imported in this arm: the corpus numbers and the query log come from ref_driver
        from types import SimpleNamespace

        from oracle import refrun

        base = reference_arm(args, corpus_name, log_name, kind, k, args.steps, args.warmup, True, batch) if refrun.driver() else None
        line = {"impl": "reference", "metric": "bm25_topk_queries_per_sec", "unit": "queries/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic"}
        if base is None:
            line["config"] = {"workload": f"{args.workload} (scale {args.scale})"}
            line["unavailable"] = "oracle/_ref/ref_driver missing (run make -C oracle ref in the build container)"
        else:
            line["config"] = {"workload": workload_text(SimpleNamespace(**refrun.corpus_spec(corpus_name, args.scale))),
                              "reference_sample": base["sample"]}
            line.update({"value": base["value"], "ms_per_step": 1e3 * batch / base["value"], "cpu_baseline": base,
                         "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(line)
        return 0

    # one process per GPU: the host worker pools of the ranks share the box's cores instead of oversubscribing them
    os.environ.setdefault("DGPU_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, world))))
    import ctypes as C

    import numpy as np
    import torch

    import diagon_b200 as dg
    from diagon_b200 import _lib

    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device: diagon_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        # torch.distributed is plumbing here: it ships the 128-byte NCCL id to the ranks and reduces the timings; the
        # search itself, its collective included, is inside libdiagon_b200.so (dgpu_sharded_*)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = _lib.load()
    spec = dg.named_corpus(corpus_name, args.scale)
    nseg = spec.num_segments
    seg_lo, seg_hi = nseg * rank // world, nseg * (rank + 1) // world
    t0 = time.time()
    reader = dg.IndexReader.synthetic(spec, local_rank, seg_lo, seg_hi)
    build_s = time.time() - t0
    if args.log2_window:
        reader.set_option("log2_window", args.log2_window)
    if args.ctas_per_sm:
        reader.set_option("ctas_per_sm", args.ctas_per_sm)
    if args.warps:
        reader.set_option("warps", args.warps)
    if args.warps_per_sm:
        reader.set_option("warps_per_sm", args.warps_per_sm)
    if args.window_docs:
        reader.set_option("window_docs", args.window_docs)
    if args.stage_log2:
        reader.set_option("stage_log2", args.stage_log2)
    if args.decode_ctas_per_sm:
        reader.set_option("decode_ctas_per_sm", args.decode_ctas_per_sm)
    if args.part_factor:
        reader.set_option("part_factor", args.part_factor)
    if args.kernel:
        reader.set_option("kernel", args.kernel)
    if args.lane_merge >= 0:
        reader.set_option("lane_merge", args.lane_merge)
    if args.union_window_docs:
        reader.set_option("union_window_docs", args.union_window_docs)
    if args.filter_stream >= 0:
        reader.set_option("filter_stream", args.filter_stream)
    if args.union_max_overlap >= 0:
        reader.set_option("union_max_overlap", args.union_max_overlap)
    if args.lane_ring_entries:
        reader.set_option("lane_ring_entries", args.lane_ring_entries)
    if args.pipeline_chunks:
        reader.set_option("pipeline_chunks", args.pipeline_chunks)
    if args.pool_smem_cap:
        reader.set_option("pool_smem_cap", args.pool_smem_cap)
    if args.lane_ctas_per_sm:
        reader.set_option("lane_ctas_per_sm", args.lane_ctas_per_sm)
    sharded = None
    if world > 1:
        # every rank joins the library's NCCL communicator; creating the sharded searcher also sums docFreq and the field
        # totals over the ranks (idf / avgdl must be the global ones on every rank)
        uid = [dg.ShardedSearcher.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sharded = dg.ShardedSearcher(reader, uid[0], rank, world)
        searcher = sharded.local
    else:
        searcher = dg.IndexSearcher(reader)
    text = dg.query_log_text(log_name, spec.vocab, batch, kind)
    nq = text.count(b"\n")
    stats = searcher.stage_batch_text(text, k)
    eng = reader.engine()
    stream = torch.cuda.Stream()          # a real (non-default) stream shared by the engine launches,
    torch.cuda.set_stream(stream)         # the NCCL collectives and the timing events
    sptr = C.c_void_p(stream.cuda_stream)

    def device_step():
        """Kernels only (plus, for N > 1, the exchange: pack, ONE ncclAllGather, merge - all inside the library, on the
        same stream). Returns nothing; async."""
        if sharded is not None:
            sharded.search_staged(stream.cuda_stream)
        elif lib.dgpu_engine_search_staged(eng, sptr) != 0:
            raise RuntimeError(lib.dgpu_engine_last_error().decode())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then exactly K steps in one timed region
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.dgpu_engine_launch_count(eng)
    kernel_ms, phase_ms = [], []
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        device_step()
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = lib.dgpu_engine_launch_count(eng) - launches0
    # per-kernel time of the search kernel alone (engine events around the launch), one more untimed step
    for _ in range(3):
        device_step()
        barrier()
        lib.dgpu_engine_sync(eng)
        kernel_ms.append(lib.dgpu_engine_last_search_ms(eng))
        ph = (C.c_float * 3)()
        lib.dgpu_engine_last_phase_ms(eng, C.byref(ph))
        phase_ms.append([float(x) for x in ph])
    bstats = (C.c_uint64 * 16)()
    lib.dgpu_engine_batch_stats(eng, C.byref(bstats))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0])
    ms_per_step = total_ms / args.steps
    value = nq / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with host buffers: query text in host memory -> results in host memory.
    # N == 1: one call, dgpu_search_batch_text. N > 1: one call per rank, dgpu_sharded_search_batch_text (the collective
    # is inside the library).
    out = searcher._alloc(nq, k)
    e2e_times = []
    n_warm = max(2, min(args.warmup, 3))
    for i in range(n_warm + args.steps):
        barrier()
        t1 = time.perf_counter()
        if sharded is None:
            res = searcher.search_batch_text(text, k, nq, out)
        else:
            # every rank hands the same batch to the library: parse + compile, H2D, kernels on its shard, one all-gather
            # per chunk, merge, D2H of the merged top-k into host arrays
            res = sharded.search_batch_text(text, k, nq, out)
        t2 = time.perf_counter()
        if i >= n_warm:
            e2e_times.append(t2 - t1)
    e2e_t = torch.tensor([sum(e2e_times)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = nq * args.steps / float(e2e_t[0])
    stream_of = searcher if sharded is None else sharded
    # a stream of batches: batch i + 1 is submitted (host work, H2D, launches) before batch i is collected (wait, D2H,
    # unpack), so the host work of one overlaps the kernels of the other and every batch runs whole. Every step still
    # carries its own text in and its own results out; the timed region holds `steps` submits and `steps` collects,
    # the pipeline's fill and drain included.
    e2e_sync = {"value": e2e_value, "ms_per_step": 1e3 * float(e2e_t[0]) / args.steps,
                "call": ("dgpu_search_batch_text" if sharded is None else "dgpu_sharded_search_batch_text on every rank") +
                        ", one call at a time (the batch is cut into 3 chunks inside the call)"}
    outs = [searcher._alloc(nq, k) for _ in range(2)]
    for timed in (False, True):
        n_steps = args.steps if timed else n_warm
        barrier()
        t1 = time.perf_counter()
        prev = stream_of.submit_batch_text(text, k)
        for i in range(1, n_steps):
            cur = stream_of.submit_batch_text(text, k)
            res = prev.collect(outs[(i - 1) & 1])
            prev = cur
        res = prev.collect(outs[(n_steps - 1) & 1])
        t2 = time.perf_counter()
    e2e_t = torch.tensor([t2 - t1], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = nq * args.steps / float(e2e_t[0])
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel: accumulate_topk_kernel (batched path) or search_kernel (fused path)
    peak, peak_src = measured_peak()
    algo = torch.tensor([stats["algorithmic_bytes"], stats["postings"]], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(algo)
    med = [statistics.median(p[i] for p in phase_ms) for i in range(3)]
    batched = (args.kernel or 3) == 3
    k_ms = med[1] if batched else statistics.median(kernel_ms)
    kt = torch.tensor([k_ms, med[0]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    achieved = stats["algorithmic_bytes"] / (float(kt[0]) / 1e3) / 1e9  # this rank's bytes / its kernel time
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu launch list (same workload only)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj["workload"] == args.workload and args.scale == 1.0 and not args.queries and world == 1 and batched \
                and tj["kernel"] == dominant_kernel(bstats):
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8000_gbs": achieved / 8000.0,   # the north star quotes the nominal peak
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": dominant_kernel(bstats) if batched else "search_kernel",
                "kernel_ms": float(kt[0]), "algorithmic_bytes_per_launch": stats["algorithmic_bytes"],
                "postings_per_launch": stats["postings"],
                "postings_per_s": stats["postings"] / (float(kt[0]) / 1e3),
                "step_ms_by_kernel": {"decode_score_kernel": med[0], "scoring_kernels": med[1], "merge_items_kernel": med[2]},
                "work_items": {"accumulate_topk_kernel": int(bstats[6]), "intersect_topk_kernel": int(bstats[7]),
                               lane_kernel_name(bstats): int(bstats[10]), "union_topk_kernel": int(bstats[13])}}
    if batched and int(bstats[10]) and int(bstats[11]) == 1:
        roofline["ring_entries_per_warp"] = int(bstats[12])
    if batched and med[0] > 0:
        # K1 alone: reads the compressed blocks of the distinct terms once, writes 8 B per decoded posting slot
        rd, wr = int(bstats[4]), int(bstats[2]) * 8
        roofline["decode_score_kernel"] = {
            "distinct_terms": int(bstats[0]), "read_bytes": rd, "write_bytes": wr, "kernel_ms": float(kt[1]),
            "achieved": (rd + wr) / (float(kt[1]) / 1e3) / 1e9, "unit": "GB/s",
            "frac": (rd + wr) / (float(kt[1]) / 1e3) / 1e9 / peak,
            "frac_of_nominal_8000_gbs": (rd + wr) / (float(kt[1]) / 1e3) / 1e9 / 8000.0}
        roofline["window_docs"] = int(bstats[5])
        roofline["doc_range_splits"] = int(bstats[3])

    if rank == 0:
        # bytes the engine copies per call, counted by the engine from the arrays it stages / fetches; the query text
        # itself (host memory) is parsed on the host and never copied
        h2d = int(bstats[8])
        d2h = int(bstats[9])
        line = {
            "metric": "bm25_topk_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(spec),
                       "sharding": f"{nseg} segments over {world} GPU(s), {'one NCCL all-gather per batch + device merge, inside the library' if world > 1 else 'single GPU'}",
                       "l2": "inputs larger than L2 (device image %.0f MB per GPU)" % (reader.image_bytes() / 1e6),
                       "index_build_s": build_s, "postings_on_gpu": reader.num_postings(), "image_bytes": reader.image_bytes(),
                       "kernel_path": "batched" if batched else "fused-windows"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "call": ("dgpu_submit_batch_text + dgpu_collect_batch, two batches in flight (host text -> host results: batch "
                             "i + 1 is parsed, compiled, staged and launched before batch i is collected)" if world == 1 else
                             "dgpu_sharded_submit_batch_text + dgpu_collect_batch on every rank, two batches in flight (host text -> "
                             "merged host results: the ranks divide parse + compile, H2D, kernels on the rank's shard, ONE "
                             "ncclAllGather of k keys + count + hits per query, device merge, D2H)"),
                    "ms_per_step": 1e3 * float(e2e_t[0]) / args.steps,
                    **({"one_call_at_a_time": e2e_sync} if e2e_sync else {})},
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:   # full-size parity against the reference itself, when its whole-corpus index is on this box
                par = parity_vs_reference(args, searcher, corpus_name, log_name, kind, k, spec)
                if par:
                    line["parity_vs_reference"] = par
            except Exception as e:
                line["parity_vs_reference"] = {"error": str(e)[:300]}
            try:
                base = reference_arm(args, corpus_name, log_name, kind, k, 1, 1, False, batch)
                if base:
                    line["cpu_baseline"] = base
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"error": str(e)[:300]}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
