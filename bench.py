#!/usr/bin/env python
"""EODM fwd+bwd throughput (frames/s) on B200, with the roofline of the dominant
kernel and the reference's CPU formulation timed beside it.

    python bench.py --gpus 1 --steps 30 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 5 --warmup 1      # the CPU arm

A step = one pass of the hot path over one synthetic batch: softmax ->
expected n-gram counts -> [all-reduce of K+1 floats] -> loss and dloss/dS ->
counts VJP -> softmax VJP (the EODM_loss boundary: `_logits` in, loss and
dloss/d_logits out; models/EODM.py:5-25 + the tape of main_EODM.py:168).
Workload at N=1: BASELINE.json configs[1] (B=256, T=400, V=48, trigram top-10k);
at N>1 every rank runs that shape on its own batch (weak scaling) and the ranks
exchange the partial counts over NCCL.  frames = valid posterior rows.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "unsupervised-asr_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

L2_FLUSH_BYTES = 256 << 20
CPU_CHUNK_B = 64          # the dense [B, T', K] intermediate of the reference graph is 1 GB per 64 utterances at timit_c2


def load_synth():
    """The synthetic-workload module WITHOUT importing the package (whose __init__ loads libeodm_b200.so): the
    reference arm must not map the product library."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "eodm_synth_standalone", os.path.join(ROOT, "unsupervised-asr_b200", "eodm_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def config_of(w, name, world):
    """The workload description both arms print: identical keys and values for the same workload."""
    return {"workload": name, "B_per_gpu": w["B"], "T": w["T"], "V": w["V"], "n": w["n"], "K": w["K"],
            "frames_per_step": int(w["mask"].sum()) * world,
            "boundary": "_logits -> loss, dloss/d_logits (softmax inside)",
            "parallelism": "batch-sharded x%d" % world,
            "l2": "GPU arm: flushed between timed steps (256 MiB write)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="timit_c2")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: how the partial counts meet -- one kernel over NVLink peer memory fused with the loss "
                         "(csrc/peer.cu), or ncclAllReduce followed by the loss kernel; auto = peer from 4 GPUs up "
                         "(measured: NCCL is 10 us ahead at N=2, the peer kernel 4 us ahead at N=8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true",
                    help="skip the strong-scaling section (BASELINE configs[2]: V=72, orders 1-5, global batch 2048)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# CPU arm: the reference formulation (softmax -> log -> dense one-hot Conv1D -> exp -> mask ->
# reduce -> loss, backward by autodiff) restated op for op on torch-CPU.  TensorFlow is not
# installable in this image, so this is kind="port" (oracle/eodm_oracle.py:eodm_loss_literal).
# ---------------------------------------------------------------------------
def cpu_reference(w, steps, warmup, sample_b=None):
    """Times the reference formulation.  sample_b=None: the WHOLE batch per step, evaluated chunk by chunk
    (CPU_CHUNK_B utterances at a time -- the graph materialises [B, T', K] several times over); otherwise the first
    sample_b utterances only (the bounded `cpu_baseline` leg of the default run)."""
    import torch

    from oracle import eodm_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = w["B"] if sample_b is None else min(sample_b, w["B"])
    kernel = O.ids_to_kernel(w["ids"], w["V"])
    chunks = [(b0, min(Bs, b0 + CPU_CHUNK_B)) for b0 in range(0, Bs, CPU_CHUNK_B)]
    frames = int(w["mask"][:Bs].sum())
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for b0, b1 in chunks:
            O.eodm_loss_literal(w["logits"][b0:b1], w["mask"][b0:b1], kernel, w["py"], dtype="float32", need_grad=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    what = ("all %d utterances of the %s batch, %d at a time" % (Bs, w["name"], CPU_CHUNK_B) if sample_b is None else
            "first %d of %d utterances of the %s batch" % (Bs, w["B"], w["name"]))
    return dict(frames=frames, times=times, cores=cores, sample_b=Bs,
                sample="%s (T=%d, V=%d, n=%d, K=%d), fp32, fwd+bwd by autograd; TF-equivalent dense restatement on "
                       "torch-CPU (TF 2.x not installable in this image)" % (what, w["T"], w["V"], w["n"], w["K"]))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth = load_synth()
    w = synth.workload(args.workload)
    w["name"] = args.workload
    r = cpu_reference(w, max(1, args.steps), max(0, args.warmup))
    sec = sum(r["times"]) / len(r["times"])
    val = r["frames"] / sec
    _emit(json.dumps({
        "impl": "reference", "metric": "EODM fwd+bwd frames/sec", "value": val, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(w, args.workload, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as td

    import eodm_b200 as E
    from eodm_b200 import synth
    from eodm_b200.session import PinnedArray

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "--gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
        comm = E.dist.Comm.from_torch_distributed()

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t[0])

    w = synth.workload(args.workload, rank=rank)
    w["name"] = args.workload
    B, T, V, n, K = w["B"], w["T"], w["V"], w["n"], w["K"]
    frames = int(w["mask"].sum())
    table = E.NgramTable.from_ids(w["ids"], V, device=local)
    sess = E.Session(table, w["py"], B, T)
    group = None
    if world > 1 and (args.exchange == "peer" or (args.exchange == "auto" and world >= 4)):
        # CUDA IPC can be unavailable (container policy): every rank must agree before the first collective step
        try:
            group = E.dist.PeerGroup.from_torch_distributed(K)
        except Exception as exc:                                        # noqa: BLE001
            print("rank %d: peer exchange unavailable (%s)" % (rank, exc), file=sys.stderr)
            group = None
        ok = torch.tensor([1 if group is not None else 0], device=dev)
        td.all_reduce(ok, op=td.ReduceOp.MIN)
        if int(ok[0]) == 1:
            sess.set_peer(group)
        else:
            if group is not None:
                group.close()
            group = None                                                # all ranks use ncclAllReduce
    logits_d = torch.tensor(w["logits"], device=dev)
    mask_d = torch.tensor(w["mask"], device=dev).to(torch.uint8)
    loss_d = torch.zeros(1, device=dev)
    dlogits_d = torch.empty_like(logits_d)
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        sess.step_device(logits_d.data_ptr(), mask_d.data_ptr(), B, T, loss_d.data_ptr(), dlogits_d.data_ptr(), stream,
                         comm=comm)

    def timed(fn, steps, warmup):
        """Per-step CUDA-event times (ms) with an L2 flush between steps."""
        for _ in range(warmup):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t0 = time.perf_counter()
        for a, b in evs:
            flush.zero_()
            a.record()
            fn()
            b.record()
        barrier()
        t1 = time.perf_counter()
        return [a.elapsed_time(b) for a, b in evs], t0, t1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    ms, t0, t1 = timed(step, args.steps, max(3, args.warmup))
    total_ms = max_over_ranks(sum(ms))
    ms_per_step = total_ms / args.steps
    frames_all = frames
    if world > 1:
        ft = torch.tensor([frames], dtype=torch.float64, device=dev)
        td.all_reduce(ft)
        frames_all = int(ft[0])
    value = frames_all / (ms_per_step * 1e-3)
    t_clk0 = t0

    # ---- N > 1: the same N batches on ONE GPU (rank 0), compared with what the ranks computed together ----
    parity = None
    if world > 1:
        step()
        torch.cuda.synchronize()
        gathered = [torch.empty_like(dlogits_d) for _ in range(world)] if rank == 0 else None
        td.gather(dlogits_d, gathered, dst=0)
        losses = [torch.empty_like(loss_d) for _ in range(world)] if rank == 0 else None
        td.gather(loss_d, losses, dst=0)
        if rank == 0:
            ws_all = [synth.workload(args.workload, rank=r) for r in range(world)]
            lg_all = torch.tensor(np.concatenate([x["logits"] for x in ws_all]), device=dev)
            mk_all = torch.tensor(np.concatenate([x["mask"] for x in ws_all]), device=dev).to(torch.uint8)
            one = E.Session(table, w["py"], B * world, T)
            l1 = torch.zeros(1, device=dev)
            d1 = torch.empty_like(lg_all)
            one.step_device(lg_all.data_ptr(), mk_all.data_ptr(), B * world, T, l1.data_ptr(), d1.data_ptr(), stream)
            torch.cuda.synchronize()
            dN = torch.cat(gathered)
            parity = {"what": "loss and dloss/d_logits of the %d sharded batches vs the same %d utterances as one batch on "
                              "rank 0's GPU" % (world, B * world),
                      "loss_rel": abs(float(loss_d) - float(l1)) / abs(float(l1)),
                      "loss_spread_over_ranks": float(max(abs(float(x) - float(loss_d)) for x in losses)),
                      "grad_max_rel": float((dN - d1).abs().max() / d1.abs().max()), "gate": 1e-6}
            parity["ok"] = bool(parity["loss_rel"] <= 1e-6 and parity["grad_max_rel"] <= 1e-6
                                and parity["loss_spread_over_ranks"] == 0.0)
            one.close()
            del lg_all, mk_all, d1, dN, gathered
        barrier()

    # ---- the two counts kernels alone (rank-local; roofline of the dominant one) ----
    px = E.softmax_fwd(logits_d)
    counts = torch.empty(K + 1, device=dev)
    gS = torch.randn(K, device=dev) * 1e-3
    ksteps = min(args.steps, 20)
    ms_f, _, _ = timed(lambda: E.counts_fwd(table, px, mask_d, out=counts), ksteps, 3)
    ms_b, _, _ = timed(lambda: E.counts_bwd(table, px, mask_d, gS), ksteps, 3)
    t_f, t_b = statistics.mean(ms_f) * 1e-3, statistics.mean(ms_b) * 1e-3
    # algorithmic flops of the direct gather-product formula: n per (window, n-gram)
    # forward, 4n-4 backward (n >= 2); windows = valid window starts
    t_idx = np.arange(T)[None, :]
    windows = int((w["mask"] & (t_idx <= T - n)).sum())
    fl_f = float(n) * windows * K
    fl_b = float(4 * n - 4 if n >= 2 else 1) * windows * K
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    fp32_peak = n_sm * 128 * 2 * sm_max * 1e6 / 1e12          # TFLOP/s, CUDA-core FMA peak at the max SM clock
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    dom_is_bwd = t_b >= t_f
    ach = (fl_b / t_b if dom_is_bwd else fl_f / t_f) / 1e12
    bwd_kernel = "eodm_tc_bwd_kernel" if E.uses_tensor_vjp(table) else "eodm_counts_bwd_kernel"
    fwd_kernel = "eodm_tc_fwd3_kernel" if E.uses_tensor_fwd(table) else "eodm_counts_fwd_kernel"
    roofline = {
        "kernel": bwd_kernel if dom_is_bwd else fwd_kernel,
        "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak,
        "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz (%s); CUDA-core FMA peak, not a tensor or HBM figure"
                       % ("MEASURED_PEAKS.json" if "sm_max_mhz" in peaks else "fallback 1965 MHz"),
        "traffic": None,
        "fwd": {"kernel": fwd_kernel, "ms": t_f * 1e3, "flops": fl_f, "tflops": fl_f / t_f / 1e12,
                "frac": fl_f / t_f / 1e12 / fp32_peak},
        "bwd": {"kernel": bwd_kernel, "ms": t_b * 1e3, "flops": fl_b, "tflops": fl_b / t_b / 1e12,
                "frac": fl_b / t_b / 1e12 / fp32_peak},
        "path_fwd_bwd": {"ms": (t_f + t_b) * 1e3, "roofline_ms": max((fl_f + fl_b) / (fp32_peak * 1e12),
                                                                    12.0 * B * T * V / (hbm_peak * 1e9)) * 1e3},
        "hbm": {"algorithmic_bytes": 12.0 * B * T * V, "peak_gbs": hbm_peak,
                "achieved_gbs": 12.0 * B * T * V / (t_f + t_b) / 1e9},
    }
    roofline["path_fwd_bwd"]["frac"] = roofline["path_fwd_bwd"]["roofline_ms"] / roofline["path_fwd_bwd"]["ms"]
    try:  # DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if args.workload == "timit_c2":
            roofline["traffic"] = traffic.get(roofline["kernel"])
    except Exception:
        pass
    # the resource that actually binds the sparse walk: one shared-memory wavefront per (trie node, 32 windows);
    # 0.90 wavefronts/clk/SM measured with tools/ubench.cu on this pool (profiles/r01_ubench.txt)
    rows32 = (B * T + 31) // 32
    wf_f, wf_b = float(table.fwd_nodes) * rows32, float(table.bwd_nodes) * rows32
    lds_peak = 0.90 * n_sm * sm_max * 1e6
    if E.uses_tensor_fwd(table):
        # the tensor-core forward is bound by the shared-memory data pipe (profiles/r02_tcfwd.md): per 128-window tile and
        # CTA (9 blocks of 256 pairs) the producers' operand loads (8 warps x 4 stages each x 12 LDS.128 x 4 wavefronts), the
        # tensor core's reads of the B tile (16 K-steps x 2 M tiles x (96 + 48) rows x 32 B / 128 B), the staging stores
        # (Pe0, Pe1, B hi / remainder: 3 x 48 KB / 128 B... measured 870 with the padding) and the staging loads (288)
        tiles = (B * T + 127) // 128 * 9
        wf_tc = float(tiles) * (1536 + 1152 + 870 + 288)
        pipe_peak = 1.0 * n_sm * sm_max * 1e6
        roofline["smem_pipe"] = {
            "fwd": {"wavefronts": wf_tc, "floor_ms": wf_tc / pipe_peak * 1e3, "frac": wf_tc / pipe_peak / t_f},
            "peak_wavefronts_per_clk_per_sm": 1.0,
            "note": "128-byte wavefronts of the shared-memory data pipe (LSU + tensor-core operand reads); ncu measures "
                    "29.5 M per launch at timit_c2 (profiles/r02_tc_ncu.txt); the resource that binds the tensor-core forward"}
    roofline["smem_wavefronts"] = {
        "fwd": {"algorithmic": wf_f, "floor_ms": wf_f / lds_peak * 1e3, "frac": wf_f / lds_peak / t_f},
        "peak_wavefronts_per_clk_per_sm": 0.90, "note": "measured LDS.32 rate; the resource that binds the trie walk"}
    if E.uses_tensor_fwd(table):
        roofline["smem_wavefronts"]["note"] += " (this table's forward runs on the tensor cores: see smem_pipe)"
    if E.uses_tensor_vjp(table):
        # the tensor-core VJP issues 2 GEMMs x ceil(VP^2/256) blocks x VP/8 K-steps x 3 (3xTF32) MMAs of 256x256x8 per pair
        # of 126-row tiles; 128 clk each on a CTA pair (tools/ubench_mma2.cu)
        vp = 16 if V <= 16 else 32 if V <= 32 else 48
        mmas = 2 * ((vp * vp + 255) // 256) * (vp // 8) * 3
        tile_pairs = ((B * T + 125) // 126 + 1) // 2
        rounds = (tile_pairs + n_sm // 2 - 1) // (n_sm // 2)
        floor_ms = rounds * mmas * 128 / (sm_max * 1e6) * 1e3
        roofline["tensor_pipe"] = {"bwd": {"mma_per_tile_pair": mmas, "tile_pairs": tile_pairs, "rounds": rounds,
                                           "floor_ms": floor_ms, "frac": floor_ms / (t_b * 1e3)},
                                   "note": "128 clk per cta_group::2 tf32 MMA at sm_max_mhz; the resource that binds the VJP"}
    else:
        roofline["smem_wavefronts"]["bwd"] = {"algorithmic": wf_b, "floor_ms": wf_b / lds_peak * 1e3,
                                              "frac": wf_b / lds_peak / t_b}

    # ---- end to end through the host-buffer session (H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        h_logits = PinnedArray((B, T, V), np.float32)
        h_mask = PinnedArray((B, T), np.uint8)
        h_dl = PinnedArray((B, T, V), np.float32)
        h_logits.array[...] = w["logits"]
        h_mask.array[...] = w["mask"]
        for _ in range(max(3, args.warmup)):
            sess.loss(h_logits.array, h_mask.array, h_dl.array, comm=comm)
        barrier()
        te0 = time.perf_counter()
        for _ in range(args.steps):
            loss_val = sess.loss(h_logits.array, h_mask.array, h_dl.array, comm=comm)
        barrier()
        e_sec = max_over_ranks(time.perf_counter() - te0) / args.steps
        e2e = {"value": frames_all / e_sec, "unit": "frames/s", "ms_per_step": e_sec * 1e3,
               "h2d_bytes_per_step": B * T * V * 4 + B * T, "d2h_bytes_per_step": B * T * V * 4 + 4,
               "loss": loss_val, "api": "eodm_session_loss (C ABI, pinned host buffers): one synchronous call per step"}
        # the same steps with two in flight (eodm_session_submit / _wait): step i+1's H2D and step i-1's D2H overlap
        # step i's compute -- what a stream of independent batches gets; reported beside, not instead of, the figure above
        bufs = [(h_logits, h_mask, h_dl, PinnedArray((1,), np.float32))]
        h2 = (PinnedArray((B, T, V), np.float32), PinnedArray((B, T), np.uint8), PinnedArray((B, T, V), np.float32),
              PinnedArray((1,), np.float32))
        h2[0].array[...] = w["logits"]
        h2[1].array[...] = w["mask"]
        bufs.append(h2)

        def run_pipelined(nsteps):
            def sub(i):
                lg, mk, dl, ls = bufs[i & 1]
                sess.submit(i & 1, lg.array, mk.array, ls.array, dl.array, comm=comm)
            sub(0)
            for i in range(1, nsteps):
                sub(i)
                sess.wait((i - 1) & 1)
            sess.wait((nsteps - 1) & 1)

        run_pipelined(max(3, args.warmup))
        barrier()
        tp0 = time.perf_counter()
        run_pipelined(args.steps)
        barrier()
        p_sec = max_over_ranks(time.perf_counter() - tp0) / args.steps
        e2e["pipelined"] = {"value": frames_all / p_sec, "unit": "frames/s", "ms_per_step": p_sec * 1e3,
                            "loss": float(bufs[(args.steps - 1) & 1][3].array[0]),
                            "api": "eodm_session_submit / eodm_session_wait, two steps in flight; same copies per step"}
    # ---- strong scaling: BASELINE configs[2] (V=72, one table per order 1-5, GLOBAL batch 2048 split over the ranks) ----
    strong = None
    if not args.no_strong:
        c3 = synth.LIBRI_C3
        if c3["B"] % world == 0:
            tabs = synth.order_tables(c3["V"], c3["orders"])
            ops3 = [E.NgramTable.from_ids(ids, c3["V"], device=local) for ids, _ in tabs]
            lg3, mk3 = synth.libri_c3_shard(rank, world)
            Bl = lg3.shape[0]
            ms3 = E.MultiOrderSession(ops3, [py_ for _, py_ in tabs], Bl, c3["T"])
            lg3_d = torch.tensor(lg3, device=dev)
            mk3_d = torch.tensor(mk3, device=dev).to(torch.uint8)
            loss3 = torch.zeros(len(tabs) + 1, device=dev)
            dl3 = torch.empty_like(lg3_d)

            def step3():
                ms3.step_device(lg3_d.data_ptr(), mk3_d.data_ptr(), Bl, c3["T"], loss3.data_ptr(), dl3.data_ptr(), stream,
                                comm=comm)

            n3 = max(3, min(args.steps, 10))
            t3, _, _ = timed(step3, n3, 3)
            ms3_step = max_over_ranks(sum(t3)) / n3
            fr3 = torch.tensor([float(mk3.sum())], dtype=torch.float64, device=dev)
            if world > 1:
                td.all_reduce(fr3)
            strong = {"workload": "libri_c3", "scaling": "strong", "global_batch": c3["B"], "B_per_gpu": Bl, "T": c3["T"],
                      "V": c3["V"], "orders": [list(o) for o in c3["orders"]], "frames_per_step": int(fr3[0]),
                      "steps": n3, "ms_per_step": ms3_step, "value": float(fr3[0]) / (ms3_step * 1e-3), "unit": "frames/s",
                      "loss": float(loss3[len(tabs)].item()),
                      "step": "one fused multi-order step (eodm_multi_step_device): one softmax, five counts, ONE "
                              "all-reduce of the packed [S_1, N, ..., S_5, N] buffer, one loss kernel, five VJPs into one "
                              "dpx, one softmax VJP"}
            if world > 1:
                # the same global batch on rank 0's GPU alone
                dl_all = [torch.empty_like(dl3) for _ in range(world)] if rank == 0 else None
                td.gather(dl3, dl_all, dst=0)
                if rank == 0:
                    lgA, mkA = synth.libri_c3_shard(0, 1)
                    one = E.MultiOrderSession(ops3, [py_ for _, py_ in tabs], c3["B"], c3["T"])
                    lgA_d = torch.tensor(lgA, device=dev)
                    mkA_d = torch.tensor(mkA, device=dev).to(torch.uint8)
                    l1 = torch.zeros(len(tabs) + 1, device=dev)
                    d1 = torch.empty_like(lgA_d)
                    one.step_device(lgA_d.data_ptr(), mkA_d.data_ptr(), c3["B"], c3["T"], l1.data_ptr(), d1.data_ptr(), stream)
                    torch.cuda.synchronize()
                    dN = torch.cat(dl_all)
                    strong["parity"] = {"loss_rel": abs(float(loss3[len(tabs)]) - float(l1[len(tabs)])) / abs(float(l1[len(tabs)])),
                                        "grad_max_rel": float((dN - d1).abs().max() / d1.abs().max()), "gate": 1e-6}
                    strong["parity"]["ok"] = bool(strong["parity"]["loss_rel"] <= 1e-6 and strong["parity"]["grad_max_rel"] <= 1e-6)
                    one.close()
                    del lgA_d, mkA_d, d1, dN, dl_all
                barrier()
            ms3.close()

    if rank == 0:
        clocks = sampler.summary(t_clk0, time.perf_counter())   # every timed region of this run
        sampler.stop()
    else:
        clocks = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(w, 3, 1, sample_b=CPU_CHUNK_B)
        best = min(r["times"])
        cpu = {"value": r["frames"] / best, "unit": "frames/s", "cores": r["cores"], "kind": "port",
               "sample": r["sample"] + "; 1 warm-up + best of 3 (%.2f s)" % best}

    if rank == 0:
        # our kernels per step.  Both counts kernels on the tensor cores: softmax, forward, tail (slice sums + loss + dloss/dS +
        # G image), VJP, softmax VJP = 5 on one GPU; with an exchange the forward's finish kernel runs before it (NCCL: 6 of
        # ours + NCCL's own; peer memory: finish, exchange+loss, G image = 7).  Trie walk: softmax, counts, finish, loss,
        # prepare_g, VJP, softmax VJP = 7.
        tc_both = E.uses_tensor_fwd(table) and E.uses_tensor_vjp(table)
        launches_per_step = 7 if not tc_both or group is not None else (5 if world == 1 else 6)
        out = {
            "metric": "EODM fwd+bwd frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(w, args.workload, world),
            "notes": {"l2": "flushed between timed steps (256 MiB write)",
                      "exchange": None if world == 1 else
                      ("one kernel over NVLink peer memory, fused with the loss" if group is not None else "ncclAllReduce"),
                      "path": ("forward: %s; slice sums, loss, dloss/dS and the VJP's G image: one launch; VJP: tcgen05 cta_group::2 "
                               "3xTF32, windows on the M axis, TMA-staged posterior tile"
                               % ("tcgen05 3xTF32, pairs on the M axis, operand formed in registers and written to TMEM"
                                  if E.uses_tensor_fwd(table) else "cuda-core trie walk (shared-memory operand tile)")
                               if E.uses_tensor_vjp(table) else
                               "cuda-core trie walk (shared-memory operand tile, flat tails, node stream two entries ahead)")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "loss": float(loss_d.item()),
        }
        if parity is not None:
            out["parity"] = parity
        if strong is not None:
            out["strong_scaling"] = strong
        _emit(json.dumps(out))
    if world > 1:
        if group is not None:
            assert not group.failed(), "a peer never arrived at the exchange"
            td.barrier()
            sess.close()
            group.close()
        comm.close()
        td.destroy_process_group()


_JSON_FD = 1


def _emit(line):
    os.write(_JSON_FD, (line + "\n").encode())


if __name__ == "__main__":
    # stdout carries exactly ONE line (the JSON record): libraries that print to fd 1 (NCCL's version banner,
    # torchrun notices) are sent to stderr instead.
    sys.stdout.flush()
    globals()["_JSON_FD"] = os.dup(1)
    os.dup2(2, 1)
    main()
