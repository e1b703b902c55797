"""Contract benchmark of the hot path: DiffMM graph propagation + fused full-sort evaluation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload scaled|baby|sports|clothing]
    python bench.py --impl reference ...      # the CPU restatement of the reference on the host cores

A STEP is one complete pass of the hot path over the synthetic workload: propagation (`forward_MM`:
projections + every SpMM, never cached across steps), fused score + train-history mask + top-K over
ALL eval users, hit matrix + Recall/NDCG/Precision/MAP.  `value` = eval users per second of the whole
job (all ranks), inputs resident in HBM; `e2e` = the same through the public Trainer/evaluator API
with the step's inputs (eval users, mask CSR, ground-truth CSR) copied from pinned host memory and
the metric vector read back inside the timed region.  One JSON line on stdout (rank 0).

Default workload: BASELINE.json configs[4] -- the 1M-user x 500k-item x 50M-interaction shape the
metric is quoted on "at 1/2/4/8 B200" (it fits one GPU; operands exceed the L2, so no flush is
needed).  Total work is fixed as N grows ("scaling": "strong").  The Baby-shaped DiffMM numbers
(configs[1]) ride along under "workloads".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOAD_NAMES = {"scaled": "diffmm_scaled_1Mx500k_50M_top50", "baby": "diffmm_baby_19445x7050_160k_top50",
                  "sports": "diffmm_sports_35598x18357_296k_top50", "clothing": "diffmm_clothing_39387x23033_278k_top50",
                  "toy": "diffmm_toy_300x120_top50"}


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of our kernels, from the committed ncu --set full
    captures (profiles/ncu_traffic.json: kernel key -> bytes); absent keys report null."""
    path = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# the GPU arm
# --------------------------------------------------------------------------------------------------


class HostInputs(object):
    """The per-step inputs of an evaluation pass in pinned host memory + their device landing buffers.  `upload()`
    enqueues the copies on a side stream; the first access to a device tensor makes the consuming stream wait for them,
    so the transfer overlaps whatever the consumer does before touching its inputs (the propagation)."""

    FIELDS = ("eval_u", "mask_rowptr", "mask_items", "gt_rowptr", "gt_items")

    def __init__(self, loader, device):
        self.host = {k: getattr(loader, k).cpu().pin_memory() for k in self.FIELDS}
        self.dev = {k: torch.empty_like(getattr(loader, k), device=device) for k in self.FIELDS}
        self.eval_len_list = loader.eval_len_list
        self.bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.stream = torch.cuda.Stream(device=device)
        self.pending = False

    def upload(self):
        self.stream.wait_stream(torch.cuda.current_stream())  # after every earlier consumer of the landing buffers
        with torch.cuda.stream(self.stream):
            for k in self.FIELDS:
                self.dev[k].copy_(self.host[k], non_blocking=True)
        self.pending = True
        return self

    def __getattr__(self, k):
        if k in HostInputs.FIELDS:
            if self.pending:
                torch.cuda.current_stream().wait_stream(self.stream)
                self.pending = False
            return self.dev[k]
        raise AttributeError(k)

    def get_eval_len_list(self):
        return self.eval_len_list


def run_gpu(args):
    import torch.distributed as dist

    from genmmrec_b200 import ops
    from genmmrec_b200.common.trainer import Trainer
    from genmmrec_b200.workload import Workload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()

    t_setup = time.time()
    wl = Workload("DiffMM", args.workload, dev, overrides={"score_precision": args.precision})
    model, cfg = wl.model, wl.config
    trainer = Trainer(cfg, model)
    k = max(cfg["topk"])
    full_loader = wl.valid
    sharded = None
    if world > 1:
        from genmmrec_b200.dist import ShardedDiffMM, shard_eval_by_user_block
        sharded = ShardedDiffMM(model)
        loader = shard_eval_by_user_block(full_loader, sharded.u0, sharded.u1)
    else:
        loader = full_loader
    n_eval_total = int(full_loader.eval_u.numel())
    host_in = HostInputs(loader, dev)
    sums_host = torch.empty((4, k), dtype=torch.float64).pin_memory()
    adj = model.norm_adj
    l2_bytes = torch.cuda.get_device_properties(dev).L2_cache_size
    operand_bytes = adj.full.algorithmic_bytes(cfg["embedding_size"])
    need_flush = operand_bytes < 2 * l2_bytes
    flush_buf = torch.empty(max(2 * l2_bytes, 1 << 28), dtype=torch.uint8, device=dev) if need_flush else None
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    def hot_path(inputs):
        """propagation (uncached) -> fused score/mask/top-K -> hits + metric sums."""
        if sharded is not None:  # row-sharded propagation (fused SpMM + all-gather), user-block sharded eval
            ue, ie = sharded.eval_factors()
            ids, _ = ops.score_mask_topk(ue.contiguous(), ie.contiguous(), k, users=inputs.eval_u,
                                         mask_rowptr=inputs.mask_rowptr, mask_items=inputs.mask_items,
                                         precision=args.precision, return_scores=False)
        else:
            model.invalidate_cache()
            ids, _ = trainer.topk_all(inputs)
        sums, _ = trainer.evaluator.metric_sums(ids, inputs)
        return sums

    def step_resident():
        sums = hot_path(loader)
        if world > 1:
            dist.all_reduce(sums)
        return sums

    def enqueue_e2e():
        sums = hot_path(host_in.upload())
        if world > 1:
            dist.all_reduce(sums)
        sums_host.copy_(sums, non_blocking=True)
        return sums_host

    def step_e2e():
        enqueue_e2e()
        torch.cuda.current_stream().synchronize()
        return sums_host

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # CUDA-graph replay of the step (with several ranks the capture holds the peer-store kernels, the second stream and
    # the NCCL barrier all-reduces; soaked at 2, 4 and 8 ranks).  GMR_GRAPH_MULTI=0 forces eager launches there.
    use_graph = (not args.no_graph) and (world == 1 or os.environ.get("GMR_GRAPH_MULTI", "1") == "1")
    graph_state = {"failed": False}

    def timed(fn, steps, warmup, profile=False, graph=False, sync_each=False):
        for _ in range(warmup):
            fn()
        barrier()
        run = fn
        if graph:
            try:
                run = trainer.graphed(fn)
                for _ in range(2):
                    run()
            except Exception as e:  # capture refused (e.g. an NCCL build without graph support): time eager launches
                sys.stderr.write("[bench] CUDA-graph capture failed (%r); timing eager launches instead\n" % (e,))
                graph_state["failed"] = True
                run = fn
                torch.cuda.synchronize()
            barrier()
        if profile:
            ops.PROFILE = []
        launches0 = ops.LAUNCHES
        evs = []
        t0 = time.perf_counter()
        for _ in range(steps):
            if flush_buf is not None:
                flush_buf.fill_(1)  # L2 flush between timed iterations (outside the per-step events)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            if sync_each:
                torch.cuda.current_stream().synchronize()  # the step's result has reached host memory
            b.record()
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, ops.LAUNCHES - launches0, prof

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # timed region: the step replayed from a CUDA graph (single GPU; eager launches otherwise).  Per-kernel durations
    # and the launch count come from a separate eager pass with CUDA events around every operator.
    ms_total, wall, launches, _ = timed(step_resident, args.steps, args.warmup, graph=use_graph)
    clocks = sampler.stop() if rank == 0 else None
    ms_eager, _, launches, prof = timed(step_resident, args.steps, 1, profile=True)
    ms_e2e, _, _, _ = timed(enqueue_e2e if use_graph else step_e2e, args.steps, max(1, args.warmup // 2), graph=use_graph,
                            sync_each=use_graph)
    if os.environ.get("GMR_PROFILE_STEP"):  # ncu --profile-from-start off: exactly one resident step is captured
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    result_dict, raw = trainer.evaluator.finalize(step_resident(), n_eval_total)

    # ---- per-kernel durations from the CUDA events recorded inside the timed region ----------------
    per = {}
    for name, meta, a, b in prof:
        key = name if name != "spmm" else "spmm[%dx%d,nnz=%d,D=%d]" % (meta["rows"], meta["cols"], meta["nnz"], meta["d"])
        e = per.setdefault(key, {"ms": [], "meta": meta, "op": name})
        e["ms"].append(a.elapsed_time(b))
    kernels = {}
    for key, e in per.items():
        avg = float(np.mean(e["ms"]))
        kk = {"avg_ms": avg, "launches_per_step": len(e["ms"]) / args.steps}
        if e["op"] == "spmm":
            kk["alg_GBs"] = e["meta"]["alg_bytes"] / avg / 1e6
            kk["gather_GBs"] = e["meta"]["nnz"] * (8 + 4 * e["meta"]["d"]) / avg / 1e6
        if e["op"] == "score_topk":
            kk["alg_TFLOPs"] = e["meta"]["flops"] / avg / 1e9
        if e["op"] == "dense_projections":
            kk["kernel"] = "gmr_dense_proj_f32 (split-TF32 tcgen05) where the shape fits, torch.mm otherwise"
            kk["TFLOPs"] = e["meta"]["flops"] / avg / 1e9
            kk["GBs"] = e["meta"]["bytes"] / avg / 1e6
        kernels[key] = kk
    step_ms = ms_total / args.steps
    for kk in kernels.values():
        kk["share_of_step"] = kk["avg_ms"] * kk["launches_per_step"] / step_ms
    spmm_keys = [kname for kname, e in per.items() if e["op"] == "spmm"]
    big_spmm = max(spmm_keys, key=lambda kname: per[kname]["meta"]["alg_bytes"])
    prop_ms = step_ms - sum(kernels[kname]["avg_ms"] * kernels[kname]["launches_per_step"]
                            for kname in kernels if per[kname]["op"] in ("score_topk", "hits_metrics"))
    traffic = load_traffic()

    def hbm_roofline(kname):
        kk = kernels[kname]
        return {"kernel": kname, "bound": "hbm", "achieved": kk["alg_GBs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": kk["alg_GBs"] / peaks["hbm_gbs"], "traffic": traffic.get(kname),
                "gather_model_GBs": kk["gather_GBs"], "avg_ms": kk["avg_ms"], "share_of_step": kk["share_of_step"],
                "algorithmic_bytes_per_launch": per[kname]["meta"]["alg_bytes"],
                "peak_source": peaks["source"] + " copy bandwidth",
                "note": "algorithmic bytes = CSR read once + X read once + Y written once (SURVEY.md 8d); gather_model_GBs "
                        "counts every gathered embedding row (what the L2 actually serves)"}

    def tensor_roofline():
        kk = kernels["score_topk"]
        return {"kernel": "score_mask_topk (%s)" % args.precision, "bound": "tensor", "achieved": kk["alg_TFLOPs"],
                "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": kk["alg_TFLOPs"] / peaks["bf16_tflops_sustained"],
                "traffic": traffic.get("score_topk"), "avg_ms": kk["avg_ms"], "share_of_step": kk["share_of_step"],
                "peak_source": peaks["source"] + " bf16 sustained (kernel timed inside a long step)",
                "note": "flops = 2*U*I*D of the reference's dense matmul.  The tc mode sweeps items in descending-norm order "
                        "and stops a 256-user group as soon as Cauchy-Schwarz rules out every remaining item, so it executes "
                        "only the tiles that can matter: on popularity-skewed embeddings `achieved` exceeds the tensor peak "
                        "because most of the dense product is provably irrelevant and never computed (exact results; "
                        "GMR_TC_DEBUG=3 forces the full sweep).  tc_split issues 3x the flops; fp32 runs on CUDA cores."}

    def proj_roofline():
        kk = kernels["dense_projections"]
        return {"kernel": "dense_proj (split-TF32 tcgen05, image + text projections)", "bound": "hbm", "achieved": kk["GBs"],
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": kk["GBs"] / peaks["hbm_gbs"], "traffic": traffic.get("dense_projections"),
                "avg_ms": kk["avg_ms"], "share_of_step": kk["share_of_step"], "tensor_TFLOPs": kk["TFLOPs"],
                "algorithmic_bytes_per_launch": per["dense_projections"]["meta"]["bytes"],
                "peak_source": peaks["source"] + " copy bandwidth",
                "note": "algorithmic bytes = feature matrices read once + outputs written once; 3 TF32 MMAs per product term"}

    # `roofline` describes the single KERNEL of ours that holds the largest share of the step.  score_topk is a
    # sequence of nine launches (operand prep, two sweeps, checkpoint, finalise, fp32 redo; the largest of them is
    # smaller than the SpMMs, profiles/r01_ncu_launches_step_*.csv) and is described by `roofline_score`.
    single = [kname for kname in kernels if per[kname]["op"] in ("spmm", "dense_projections")]
    dominant = max(single, key=lambda kname: kernels[kname]["avg_ms"])
    roofline = proj_roofline() if per[dominant]["op"] == "dense_projections" else hbm_roofline(dominant)
    roofline_spmm = hbm_roofline(big_spmm)
    roofline_score = tensor_roofline()

    out = None
    if rank == 0:
        out = {
            "metric": "full_sort_eval_users_per_s", "value": n_eval_total / step_ms * 1e3, "unit": "users/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[args.workload], "propagation": "DiffMM forward_MM (GenMMRec/src/models/diffmm.py:129-169)", "n_users": wl.n_users,
                       "n_items": wl.n_items, "nnz_train": wl.nnz_train, "eval_users": n_eval_total, "topk": k,
                       "embedding_size": cfg["embedding_size"], "n_layers": cfg["n_layers"],
                       "score_precision": args.precision,
                       "launch": "CUDA graph replay of the whole step" if (use_graph and not graph_state["failed"]) else "eager launches",
                       "parallelism": ("row-sharded propagation (push-SpMM all-gather) + user-block sharded eval, x%d" % world)
                       if world > 1 else "single GPU",
                       "l2": ("L2 flushed between timed steps (operands fit the %d MB L2)" % (l2_bytes >> 20)) if need_flush
                       else ("operands (%.0f MB per SpMM) exceed the %d MB L2; no flush" % (operand_bytes / 1e6, l2_bytes >> 20))},
            "propagation_step_ms": prop_ms, "spmm_hbm_GBs": kernels[big_spmm]["alg_GBs"],
            "roofline": roofline, "roofline_spmm": roofline_spmm, "roofline_score": roofline_score,
            "roofline_proj": proj_roofline() if "dense_projections" in kernels else None, "kernels": kernels,
            "e2e": {"value": n_eval_total / (ms_e2e / args.steps) * 1e3, "unit": "users/s",
                    "h2d_bytes_per_step": host_in.bytes * world, "d2h_bytes_per_step": int(sums_host.numel() * 8),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "eager_ms_per_step": ms_eager / args.steps, "clocks": clocks, "wall_s_timed_region": wall, "setup_s": setup_s,
            "metrics": result_dict,
        }
    return out, wl, trainer, full_loader


# --------------------------------------------------------------------------------------------------
# the CPU arm: oracle port of the reference, timed on the host cores
# --------------------------------------------------------------------------------------------------


class CpuReference(object):
    """The reference's own evaluation step restated on CPU tensors (oracle/ref_port.py): per batch of
    `eval_batch_size` users it re-runs forward_MM (models/diffmm.py:276), multiplies, masks, takes
    torch.topk and runs the Python hit loop + numpy metrics (common/trainer.py:379-388,
    utils/topk_evaluator.py:107-120)."""

    def __init__(self, wl, loader, batch_users):
        from oracle import ref_port as rp
        self.rp = rp
        m = wl.model
        self.cfg = wl.config
        self.nu = wl.n_users

        def coo(g):
            t = g.to_torch_coo()
            return torch.sparse_coo_tensor(t._indices().cpu(), t._values().cpu(), t.shape)  # uncoalesced, as shipped

        self.adj = coo(m.norm_adj.full)
        self.img_adj, self.txt_adj = coo(m.image_UI_matrix), coo(m.text_UI_matrix)
        self.p = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        self.v_feat, self.t_feat = m.v_feat.cpu(), m.t_feat.cpu()
        b = min(batch_users, int(loader.eval_u.numel()))
        self.b = b
        self.users = loader.eval_u[:b].cpu()
        rp_, it_ = loader.mask_rowptr[:b + 1].cpu(), loader.mask_items.cpu()
        rows = torch.repeat_interleave(torch.arange(b), rp_[1:] - rp_[:-1])
        self.mask = torch.stack([rows, it_[: int(rp_[-1])].long()])
        g_rp = loader.gt_rowptr[:b + 1].cpu().numpy()
        g_it = loader.gt_items.cpu().numpy()
        self.gt = [g_it[g_rp[j]:g_rp[j + 1]] for j in range(b)]
        self.gt_len = np.diff(g_rp)
        self.k = max(self.cfg["topk"])
        self.parts = {}

    def step(self):
        rp, cfg = self.rp, self.cfg
        t0 = time.perf_counter()
        with torch.no_grad():
            ue, ie = rp.diffmm_forward_mm(self.p, self.adj, self.img_adj, self.txt_adj, self.v_feat, self.t_feat, self.nu,
                                          cfg["n_layers"], cfg["ris_lambda"], cfg["ris_adj_lambda"])
            t1 = time.perf_counter()
            scores = torch.matmul(ue[self.users], ie.transpose(0, 1))
            scores[self.mask[0], self.mask[1]] = -1e10
            _, idx = torch.topk(scores, self.k, dim=-1)
        t2 = time.perf_counter()
        hit = rp.hit_matrix(idx.numpy(), self.gt)
        raw = np.stack([rp.METRICS[mm.lower()](hit, self.gt_len) for mm in cfg["metrics"]], axis=0)
        t3 = time.perf_counter()
        self.parts = {"forward_MM_s": t1 - t0, "score_mask_topk_s": t2 - t1, "hits_metrics_s": t3 - t2}
        return t3 - t0, raw


def cpu_baseline(wl, loader, steps=1, warmup=0, batch_users=4096):
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    ref = CpuReference(wl, loader, batch_users)
    for _ in range(warmup):
        ref.step()
    ts = [ref.step()[0] for _ in range(steps)]
    t = float(np.mean(ts))
    return {"value": ref.b / t, "unit": "users/s", "cores": cores, "kind": "port",
            "sample": "one reference evaluation batch of %d users (incl. the forward_MM the reference re-runs per "
                      "batch, the [B, n_items] matmul, mask, torch.topk and the Python hit loop + numpy metrics); "
                      "users/s = batch / batch time, i.e. linear extrapolation over the %d eval users"
                      % (ref.b, int(loader.eval_u.numel())),
            "seconds_per_batch": t, "parts": ref.parts}, ts


def run_reference(args):
    """--impl reference: the CPU restatement of the reference on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return None
    from genmmrec_b200.workload import Workload

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    wl = Workload("DiffMM", args.workload, dev, overrides={"score_precision": "fp32"})  # same synthetic inputs
    budget = 240.0
    base, _ = cpu_baseline(wl, wl.valid, steps=1, warmup=0)
    t1 = base["seconds_per_batch"]
    steps = max(1, min(args.steps, int(budget / t1) - 1))
    warm = 1 if args.warmup > 0 and (steps + 1) * t1 < budget else 0
    base, ts = cpu_baseline(wl, wl.valid, steps=steps, warmup=warm)
    k = max(wl.config["topk"])
    return {
        "impl": "reference", "metric": "full_sort_eval_users_per_s", "value": base["value"], "unit": "users/s",
        "n_gpus": world, "steps": args.steps, "steps_run": steps, "warmup": args.warmup, "warmup_run": warm + 1,
        "ms_per_step": base["seconds_per_batch"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload], "propagation": "DiffMM forward_MM (GenMMRec/src/models/diffmm.py:129-169)", "n_users": wl.n_users,
                   "n_items": wl.n_items, "nnz_train": wl.nnz_train, "eval_users": wl.n_eval_users, "topk": k,
                   "embedding_size": wl.config["embedding_size"], "n_layers": wl.config["n_layers"]},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def _claim_stdout():
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr and return a
    writer on the real stdout, so that the ONE JSON line is the only thing the driver reads there."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out_stream = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="scaled", choices=sorted(WORKLOAD_NAMES))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("GMR_SCORE_PRECISION", "tc"), choices=["fp32", "tc", "tc_split"],
                    help="scoring path: tc = fp16 tcgen05 certified screen + exact fp32 re-score (default), tc_split = "
                         "split-bf16 tcgen05 + exact re-rank, fp32 = CUDA cores; all three return the same ids/scores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay of the step")
    ap.add_argument("--no-extra-workloads", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")

    if args.impl == "reference":
        out = run_reference(args)
        if out is not None:
            out_stream.write(json.dumps(out) + "\n")
            out_stream.flush()
        return

    out, wl, trainer, loader = run_gpu(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            base, _ = cpu_baseline(wl, loader, steps=1, warmup=0)
            out["cpu_baseline"] = base
        if world == 1 and not args.no_extra_workloads and args.workload != "baby":
            del wl, trainer, loader
            torch.cuda.empty_cache()
            sub = argparse.Namespace(**vars(args))
            sub.workload, sub.steps, sub.warmup = "baby", 20, 5
            o2, wl2, _, ld2 = run_gpu(sub)
            extra = {k: o2[k] for k in ("value", "unit", "ms_per_step", "propagation_step_ms", "spmm_hbm_GBs", "e2e",
                                        "gpu_launches", "metrics")}
            extra["config"] = o2["config"]
            extra["roofline_spmm"] = o2["roofline_spmm"]
            if not args.no_cpu_baseline:
                extra["cpu_baseline"], _ = cpu_baseline(wl2, ld2, steps=1, warmup=1)
            out["workloads"] = {WORKLOAD_NAMES["baby"]: extra}
        out_stream.write(json.dumps(out) + "\n")
        out_stream.flush()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
