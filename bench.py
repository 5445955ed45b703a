#!/usr/bin/env python
"""Headline benchmark: QPS at recall@10 >= 0.9 of the Neural-LSH query path on B200.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this repo (CUDA hot path)
    python bench.py --impl reference --gpus 1 --steps 3 ...   # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the one the metric is quoted on; it fits one GPU):
synthetic 10M x 128 fp32 L2 database, 4096 buckets (hash_size 12, MultiLayerRelu 256,256
hasher fitted on the synthetic mixture), 10k queries per step, k = 10, multi-probe p = the
smallest power of two reaching recall@10 >= 0.9 against brute-force ground truth.  With N
ranks the SAME database is row-sharded (strong scaling): every rank scans its shard and the
per-shard top-k lists are merged after one NCCL all-gather.  A step = one pass of the 10k
query batch through hash -> probe selection -> candidate scan + top-k -> (all-gather, merge).
Prints ONE JSON line (see the key list in DESIGN.md §Measurement).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "neural-locality-sensitive-hashing_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: rows, dim, hash_size, queries, k, metric, seed   
    "cfg1_100k_x128_16b": (100_000, 128, 4, 1_000, 10, "l2", 1001),
    "cfg2_1m_x128_256b": (1_000_000, 128, 8, 10_000, 10, "l2", 1002),
    "cfg3_1.2m_x100_1024b_angular": (1_200_000, 100, 10, 10_000, 10, "angular", 1003),
    "cfg4_10m_x128_4096b": (10_000_000, 128, 12, 10_000, 10, "l2", 1004),
    "cfg5_1m_x960_512b_k100": (1_000_000, 960, 9, 1_000, 100, "l2", 1005),
}
DEFAULT_WORKLOAD = "cfg4_10m_x128_4096b"
RECALL_TARGET = 0.9
PROBE_LADDER = (1, 2, 4, 8, 16, 32, 64, 128)
SEP = 3.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the row count (debug)")
    ap.add_argument("--queries", type=int, default=0, help="override the query count (debug)")
    ap.add_argument("--probes", type=int, default=0, help="fix p instead of searching the ladder")
    ap.add_argument("--fit-steps", type=int, default=300)
    ap.add_argument("--cpu-sample", type=int, default=500, help="queries in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the CUDA graph")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
def metric_distance(metric):
    if metric == "l2":
        return F.pairwise_distance
    fn = lambda a, b: 1 - F.cosine_similarity(a, b, dim=-1)  # noqa: E731
    fn.nlsh_metric = "angular"
    return fn


def make_hashing(d, hs, metric, seed, device, fit_steps, src_rank_trains=True):
    """MultiLayerRelu(d, [256, 256]) + Linear(256, hs) + sigmoid (main.py:279-284 defaults),
    fitted on rank 0 and broadcast so every rank hashes with identical weights."""
    import synth
    from encoders import MultiLayerRelu
    from nlsh.hashings import MultivariateBernoulli
    torch.manual_seed(seed)
    hashing = MultivariateBernoulli(MultiLayerRelu(d, [256, 256]), hs, metric_distance(metric))
    hashing._hasher.to(device)
    rank = dist.get_rank() if dist.is_initialized() else 0
    loss = 0.0
    if rank == 0 or not src_rank_trains:
        loss = synth.fit_hasher(hashing, d, hs, seed, device, sep=SEP, steps=fit_steps)
    if dist.is_initialized() and dist.get_world_size() > 1:
        for prm in hashing.parameters():
            dist.broadcast(prm.data, src=0)
    hashing.train_mode(False)
    return hashing, loss


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the bench is under load."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100", "-f", self.path],
                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def count(self):
        """Samples written so far."""
        try:
            return sum(1 for line in open(self.path) if line.strip())
        except (OSError, TypeError):
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8 or not parts[0].isdigit() or int(parts[0]) != self.gpu_index:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def timed_steps(step_fn, steps, barrier):
    """EXACTLY `steps` calls bracketed by barrier + synchronize, timed with CUDA events on the
    launching stream; returns this rank's elapsed ms."""
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_fn()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1)


def max_over_ranks(value, device):
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def bucket_recall(gt_ids, probes, codes_of_gt):
    """recall@k of a bucket-probing search = fraction of the true neighbours whose bucket is
    probed (every probed bucket is scanned exhaustively, so a true neighbour that is a
    candidate is always returned)."""
    hit = (codes_of_gt[:, :, None] == probes[:, None, :]).any(-1)
    return float(hit.float().mean().item())


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def cpu_query_baseline(oracle, layers, X_host, index2row, Q_host, probe_sets, metric, k, repeats=1):
    """The reference's query path (Indexer.query, indexer.py:56-96, incl. query hashing) on
    the host cores: QPS = n / wall, as nlsh/trainers/base.py:93-107 defines it."""
    cpu = oracle.CpuIndexer.__new__(oracle.CpuIndexer)
    cpu.layers, cpu.head, cpu.metric = layers, oracle.HEAD_SIGMOID, metric
    cpu.X = X_host
    cpu._hash_codes = oracle.hash_codes_c if os.path.exists(os.path.join(ROOT, "oracle", "liboracle_pack.so")) \
        else oracle.hash_codes
    cpu.index2row = index2row
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu.hash(Q_host)  # query hashing is inside the reference's timed region
        ids, ncand = cpu.query(Q_host, k=k, probe_sets=probe_sets)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return Q_host.shape[0] / best, ids, ncand


def index2row_from_csr(offsets_host, ids_host):
    ids_t = torch.from_numpy(ids_host.astype(np.int64))
    return {int(c): ids_t[int(offsets_host[c]):int(offsets_host[c + 1])]
            for c in range(len(offsets_host) - 1) if offsets_host[c + 1] > offsets_host[c]}


# ------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import synth
    import precompute
    from nlsh import _native
    from nlsh.metrics import recall_at_k_tensors
    from nlsh.parallel import ShardedIndexer, all_gather_topk, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    n, d, hs, nq, k, metric, seed = WORKLOADS[args.workload]
    n = args.rows or n
    nq = args.queries or nq
    lo, hi = shard_range(n, rank, world)
    X = synth.make_database(n, d, hs, seed, device, sep=SEP, row_lo=lo, row_hi=hi)
    Q = synth.make_queries(nq, d, hs, seed, device, sep=SEP)
    hashing, fit_loss = make_hashing(d, hs, metric, seed, device, args.fit_steps)

    t0 = time.perf_counter()
    index = ShardedIndexer(hashing, X, hashing.distance, shard_lo=lo, metric=metric)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0

    # ---- ground truth (brute force, same sharding) and the probe count for recall >= 0.9 ----
    gt_ids, gt_d = precompute.knn_tensors(Q, X, metric, k, id_offset=lo)
    if world > 1:
        g_ids, g_d = all_gather_topk(gt_ids, gt_d)
        gt_ids, gt_d = _native.merge_topk(g_d, g_ids)
    recalls = {}
    p_used = args.probes
    ladder = (args.probes,) if args.probes else tuple(p for p in PROBE_LADDER if p <= (1 << hs))
    for p in ladder:
        ids, _, _ = index.query_tensors(Q, k=k, hash_times=p)
        recalls[p] = recall_at_k_tensors(gt_ids, ids)
        p_used = p
        if recalls[p] >= RECALL_TARGET:
            break
    recall = recalls[p_used]

    # the query step as the public API offers it for a fixed batch shape: captured once into a
    # CUDA graph (ShardedIndexer.capture_query), replayed per step
    run_query = None
    if not args.no_graph:
        try:
            run_query = index.capture_query(nq, k=k, hash_times=p_used)
        except Exception as exc:  # noqa: BLE001 - report and measure the eager path instead
            print(f"[bench] CUDA-graph capture failed ({exc!r}); timing eager launches", file=sys.stderr)
    if run_query is None:
        run_query = lambda qv: index.query_tensors(qv, k=k, hash_times=p_used)  # noqa: E731

    def step():
        return run_query(Q)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    launches0 = _native.kernel_launch_count()
    ms = max_over_ranks(timed_steps(step, args.steps, barrier), device)
    launches = _native.kernel_launch_count() - launches0
    if hasattr(run_query, "kernels_per_call"):  # graph replays do not pass through the counter
        launches = run_query.kernels_per_call * args.steps

    # ---- end to end: pinned host queries in, pinned host results out, every step -------------
    # through the package's serving loop (nlsh.parallel.PipelinedSearch): two batches in flight, the
    # H2D copy of the next batch overlaps the search of the current one, every result is read on the host
    from nlsh.parallel import PipelinedSearch
    q_pinned = Q.cpu().pin_memory()
    e2e_api = "nlsh.parallel.PipelinedSearch (2 batches in flight) over ShardedIndexer.capture_query, pinned host in/out"
    try:
        pipe = PipelinedSearch(index, nq, k=k, hash_times=p_used)
        pending = []

        def e2e_step():
            pending.append(pipe.submit(q_pinned))
            if len(pending) > 1:
                ids_h, _, _ = pipe.result(pending.pop(0))  # the caller reads every result (one step late)
                assert ids_h.shape[0] == nq

        def e2e_drain():
            while pending:
                pipe.result(pending.pop(0))
    except Exception as exc:  # noqa: BLE001 - fall back to the unpipelined loop and say so
        print(f"[bench] PipelinedSearch unavailable ({exc!r}); timing the serial loop", file=sys.stderr)
        e2e_api = "ShardedIndexer query (serial), pinned host in/out"
        out_ids = torch.empty((nq, k), dtype=torch.int64).pin_memory()
        out_d = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        out_n = torch.empty((nq,), dtype=torch.int32).pin_memory()
        q_dev = torch.empty_like(Q)

        def e2e_step():
            q_dev.copy_(q_pinned, non_blocking=True)
            ids, dd, nc = run_query(q_dev)
            out_ids.copy_(ids, non_blocking=True)
            out_d.copy_(dd, non_blocking=True)
            out_n.copy_(nc, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        def e2e_drain():
            pass

    for _ in range(3):
        e2e_step()
    e2e_drain()
    torch.cuda.synchronize()

    def e2e_timed():
        e2e_step()

    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        e2e_timed()
    e2e_drain()  # the last results are read inside the timed region
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1), device)

    # ---- scan-kernel roofline: CUDA events around the kernel inside the library --------------
    _native.profile_enable(True)
    prof_steps = min(max(args.steps, 3), 200)
    torch.cuda.nvtx.range_push("nlsh_steps")  # ncu --nvtx --nvtx-include "nlsh_steps/": one step's launches
    for _ in range(prof_steps):
        _, _, ncand_local = index.local.query_tensors(Q, k=k, hash_times=p_used)
    torch.cuda.nvtx.range_pop()
    scan_ms = _native.profile_read()
    # the same kernel where no bucket tile is shared between queries (about one probing query
    # per two buckets): every candidate byte has to come from HBM, so this is the figure to
    # hold against the HBM roofline
    n_lr = max(8, min(nq, (1 << hs) // (2 * p_used)))
    for _ in range(3):
        index.local.query_tensors(Q[:n_lr], k=k, hash_times=p_used)
    _native.profile_read()
    lr_steps = 50
    for _ in range(lr_steps):
        _, _, ncand_lr = index.local.query_tensors(Q[:n_lr], k=k, hash_times=p_used)
    lr_ms = _native.profile_read()
    _native.profile_enable(False)
    # nvidia-smi needs about a second to deliver its first sample; short runs (a 0.8 ms step x 50)
    # would end before it: keep the same query step running, untimed, until a few samples exist
    # (a fixed number of steps derived from the rank-maximum step time, so that every rank issues the
    # same number of all-gathers)
    if sampler.proc is not None:
        for _ in range(int(min(20000, max(0.0, 1500.0 / max(ms / args.steps, 1e-3))))):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    lr_bytes = float(ncand_lr.double().sum().item()) * (4 * d + 4) + n_lr * (4 * d + 8 * k)
    # bytes of the DISTINCT buckets those queries probe: what has to come from HBM at least once
    lr_probes = index.local.hash_tensors(Q[:n_lr], p_used)
    lr_sizes = torch.from_numpy(index.local.bucket_sizes).to(device)
    lr_distinct = float(lr_sizes[torch.unique(lr_probes[lr_probes >= 0]).long()].sum().item()) * (4 * d + 4)
    lr_avg_ms = float(np.mean(lr_ms)) if lr_ms else float("nan")
    algo_bytes = float(ncand_local.double().sum().item()) * (4 * d + 4) + nq * (4 * d + 8 * k)
    scan_avg_ms = float(np.mean(scan_ms)) if scan_ms else float("nan")
    peaks, peak_kind = load_peaks()
    achieved = algo_bytes / (scan_avg_ms * 1e-3) / 1e9
    mean_cand = float(ncand_local.double().mean().item())

    qps = nq * args.steps / (ms * 1e-3)
    e2e_qps = nq * args.steps / (e2e_ms * 1e-3)
    scan_impl = _native.scan_impl(d, k, index.local._metric, index.local._x_sqnorm is not None, nq, p_used, 1 << hs)
    traffic = None
    prof_path = os.path.join(ROOT, "profiles", "ncu_scan_traffic.json")
    if os.path.exists(prof_path):
        rec = json.load(open(prof_path)).get(
            f"{args.workload}/p{p_used}/gpus{world}/{'tc' if scan_impl == 1 else 'simt'}")
        traffic = rec["dram_bytes_per_launch"] if rec else None
    # bytes of the DISTINCT buckets the batch probes: what HBM has to deliver at least once per step
    all_probes = index.local.hash_tensors(Q, p_used)
    distinct = float(lr_sizes[torch.unique(all_probes[all_probes >= 0]).long()].sum().item()) * (4 * d + 4)
    scan_kernel_name = ("scan_tc_kernel (tcgen05 tf32 filter + exact fp32 re-rank + top-k)" if scan_impl == 1
                        else "scan_kernel (fp32 SIMT candidate scan + top-k)")
    line = {
        "metric": "QPS at recall@10>=0.9", "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "b200",
        "config": {
            "workload": args.workload, "rows": n, "dim": d, "buckets": 1 << hs, "queries_per_step": nq,
            "k": k, "metric": metric, "probes": p_used, "recall_at_k": recall,
            "recall_by_probes": {str(a): b for a, b in recalls.items()},
            "mean_candidates_per_query_per_rank": mean_cand, "parallelism": f"row-shard x{world}",
            "hasher": "MultiLayerRelu(256,256)+sigmoid, fitted %d Adam steps (loss %.4f)" % (args.fit_steps, fit_loss),
            "l2_policy": "database shard (%.0f MB) larger than the 126 MB L2" % (X.numel() * 4 / 1e6),
            "index_build_s": build_s,
        },
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4,
                "d2h_bytes_per_step": nq * k * 12 + nq * 4, "api": e2e_api},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": scan_kernel_name, "achieved": achieved,
                     "peak": peaks["hbm_gbs"], "peak_kind": peak_kind, "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": scan_avg_ms,
                     "kernel_share_of_step": scan_avg_ms / (ms / args.steps),
                     "distinct_bucket_bytes": distinct,
                     "achieved_distinct": distinct / (scan_avg_ms * 1e-3) / 1e9,
                     "frac_distinct": distinct / (scan_avg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "dram_gbs": (traffic / (scan_avg_ms * 1e-3) / 1e9) if traffic else None,
                     "note": "algorithmic bytes count every (query, candidate) pair; a bucket tile "
                             "shared by several queries is fetched once, so achieved may exceed peak; "
                             "*_distinct counts every probed bucket once (the bytes HBM must deliver), "
                             "dram_gbs is ncu's dram__bytes per launch over the live kernel time",
                     "fp32_lane_ops_frac": (float(ncand_local.double().sum().item()) * d * (3 if metric == "l2" else 1))
                     / (scan_avg_ms * 1e-3) / (torch.cuda.get_device_properties(device).multi_processor_count
                                               * 128 * clocks.get("sm_max_mhz", 1965.0) * 1e6
                                               if clocks.get("sm_max_mhz") else 148 * 128 * 1.965e9),
                     "hbm_bound_case": {"queries": n_lr, "probes": p_used, "kernel_ms": lr_avg_ms,
                                        "kernel": "scan_tc_kernel" if _native.scan_impl(
                                            d, k, index.local._metric, index.local._x_sqnorm is not None, n_lr,
                                            p_used, 1 << hs) == 1 else "scan_kernel (fp32 SIMT)",
                                        "algorithmic_bytes_per_launch": lr_bytes,
                                        "achieved": lr_bytes / (lr_avg_ms * 1e-3) / 1e9,
                                        "frac": lr_bytes / (lr_avg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                        "distinct_bucket_bytes": lr_distinct,
                                        "achieved_distinct": lr_distinct / (lr_avg_ms * 1e-3) / 1e9,
                                        "frac_distinct": lr_distinct / (lr_avg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                        "note": "same kernel with so few queries that buckets are (almost) "
                                                "never shared; *_distinct counts each probed bucket once = "
                                                "the bytes HBM must deliver"}},
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import nlsh_oracle as oracle
        from nlsh.hashings import extract_layer_tensors
        layers = [oracle.Layer(w.cpu(), None if b is None else b.cpu(), act == 1)
                  for w, b, act in extract_layer_tensors(hashing._hasher._encoder, hashing._hasher.output_layer)]
        ns = min(args.cpu_sample, nq)
        probes = index.local.hash_tensors(Q[:ns], p_used).cpu().numpy()
        sets = [set(int(c) for c in row if c >= 0) for row in probes]
        index2row = index2row_from_csr(index.local._offsets_host, index.local._ids.cpu().numpy())
        cpu_qps, c_ids, _ = cpu_query_baseline(oracle, layers, X.cpu(), index2row, Q[:ns].cpu(), sets, metric, k)
        ids_chk, _, _ = index.local.query_tensors(Q[:ns], k=k, hash_times=p_used)
        agree = float(np.mean([a == b for a, b in zip(ids_chk.cpu().tolist(), c_ids)]))
        line["cpu_baseline"] = {
            "value": cpu_qps, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first {ns} of the {nq} queries against the full {n}-row index, same hasher and "
                      f"probe sets (p={p_used}); includes query hashing; index content taken from the GPU build",
            "host_cpus": os.cpu_count(), "ids_identical_to_gpu": agree}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle port), host cores
# ------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth
    from oracle import nlsh_oracle as oracle
    from nlsh.hashings import extract_layer_tensors

    n, d, hs, nq, k, metric, seed = WORKLOADS[args.workload]
    n = args.rows or n
    nq = args.queries or nq
    # set-up only (data, hasher fit, bucket codes, ground truth) may use torch on the GPU when
    # there is one; none of this repo's kernels run, and everything timed is CPU
    setup_dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    X = synth.make_database(n, d, hs, seed, setup_dev, sep=SEP)
    ns = min(args.cpu_sample, nq)
    Q = synth.make_queries(nq, d, hs, seed, setup_dev, sep=SEP)[:ns]
    hashing, fit_loss = make_hashing(d, hs, metric, seed, setup_dev, args.fit_steps)
    with torch.no_grad():
        enc, out = hashing._hasher._encoder, hashing._hasher.output_layer
        codes = torch.cat([((torch.sigmoid(out(enc(X[s:s + 262144]))) > 0.5).long() *
                            (1 << torch.arange(hs - 1, -1, -1, device=setup_dev))).sum(1)
                           for s in range(0, n, 262144)])
        q_logits = out(enc(Q)).cpu().numpy()
        order = torch.sort(codes, stable=True)[1]
        counts = torch.bincount(codes, minlength=1 << hs).cpu().numpy()
        offsets = np.concatenate([[0], np.cumsum(counts)])
        # ground truth for the sample (torch matmul), to pick p as the other arm does
        gt = []
        for s in range(0, ns, 128):
            dmat = torch.cdist(Q[s:s + 128], X) if metric == "l2" else \
                1 - F.normalize(Q[s:s + 128], dim=1) @ F.normalize(X, dim=1).T
            gt.append(dmat.topk(k, dim=1, largest=False)[1])
        gt = torch.cat(gt)
        codes_of_gt = codes[gt].cpu()
    index2row = index2row_from_csr(offsets, order.cpu().numpy())
    p_used, recalls = args.probes, {}
    ladder = (args.probes,) if args.probes else tuple(p for p in PROBE_LADDER if p <= (1 << hs))
    for p in ladder:
        probes = oracle.topp_probes(q_logits, oracle.HEAD_SIGMOID, p)
        recalls[p] = bucket_recall(gt.cpu(), torch.from_numpy(probes).long(), codes_of_gt)
        p_used = p
        if recalls[p] >= RECALL_TARGET:
            break
    sets = [set(int(c) for c in row if c >= 0) for row in probes]
    layers = [oracle.Layer(w.cpu(), None if b is None else b.cpu(), act == 1)
              for w, b, act in extract_layer_tensors(enc, out)]
    X_host, Q_host = X.cpu(), Q.cpu()
    del X
    for _ in range(min(args.warmup, 1)):
        cpu_query_baseline(oracle, layers, X_host, index2row, Q_host[:32], sets[:32], metric, k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        qps_step, _, ncand = cpu_query_baseline(oracle, layers, X_host, index2row, Q_host, sets, metric, k)
    wall = time.perf_counter() - t0
    qps = ns * args.steps / wall
    cores = torch.get_num_threads()
    sample = (f"each step = {ns} of the {nq} queries against the full {n}-row index on the host cores "
              f"(p={p_used}, includes query hashing); set-up (data, hasher fit, bucket grouping) untimed")
    line = {
        "metric": "QPS at recall@10>=0.9", "value": qps, "unit": "queries/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": args.workload, "rows": n, "dim": d, "buckets": 1 << hs,
                   "queries_per_step": ns, "k": k, "metric": metric, "probes": p_used,
                   "recall_at_k": recalls[p_used], "mean_candidates_per_query": float(np.mean(ncand)),
                   "hasher": "MultiLayerRelu(256,256)+sigmoid, fitted %d Adam steps (loss %.4f)" % (args.fit_steps, fit_loss)},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    _args = parse_args()
    if _args.impl == "reference":
        run_reference(_args)
    else:
        run_b200(_args)
